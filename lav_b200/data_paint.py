"""Offline point painter: paints every LiDAR sweep of a recorded dataset with the RGB segmentation model and writes the painted
features back into the dataset — the `lav.data_paint` stage of the reference's training pipeline (lav/data_paint.py:44-107,
lav/utils/datasets/point_paint_dataset.py:8-46, docs/TRAINING.md), on the lav_b200 kernels.

On-disk format = the reference's (lav/utils/datasets/basic_dataset.py:52-53,82-101 and point_paint_dataset.py:34-46): one
key-value environment per trajectory with
    len              ascii int       number of frames            town             ascii
    lidar_%05d       float32 n x 4   (x, y, z, intensity)        rgb_{cam}_%05d   PNG / JPEG bytes, BGR as cv2 writes them
    lidar_sem_%05d   float32 n x C   <- written here: softmax(seg)[1:] * (1 - softmax(seg)[0]) gathered at each point's pixel
The environment is LMDB when the `lmdb` package is importable (the reference's Dockerfile installs it; this image does not have
it) and otherwise a directory of one file per key with the SAME keys, so the painter and its tests run everywhere.

Reference loop: one frame at a time — seg model on 3 images, softmax to the host, fp64 numpy projection + gather, one LMDB
transaction per frame, 4 Ray actors per GPU.  Here: `frames_per_batch` frames per launch set — ERFNet on 3F images, then the
fused gather (ops.paint_deconv_batched: output layer + softmax + background suppression evaluated for the hit pixels only) on
the NaN-padded (F, Nmax, 4) sweep buffer; one D2H per batch.
    python -m lav_b200.data_paint --data-dir DIR --seg-weights seg_1.th [--frames-per-batch 32]
"""
import argparse
import glob
import os

import numpy as np
import torch

from . import ops
from . import point_painting as PP

try:
    import lmdb
except ImportError:          # this image: the directory store below keeps the key layout
    lmdb = None


class DirEnv:
    """one-file-per-key stand-in for an LMDB environment (same keys, same bytes)."""

    def __init__(self, path):
        self.path = path
        os.makedirs(os.path.join(path, "kv"), exist_ok=True)

    def get(self, key):
        fn = os.path.join(self.path, "kv", key)
        if not os.path.exists(fn):
            return None
        with open(fn, "rb") as f:
            return f.read()

    def put(self, key, value):
        with open(os.path.join(self.path, "kv", key), "wb") as f:
            f.write(bytes(value))

    def close(self):
        pass


class LmdbEnv:
    def __init__(self, path, write=False):
        self.env = lmdb.open(path, map_size=int(1e10)) if write else lmdb.open(path, readonly=True, lock=False, readahead=False, meminit=False)
        self.write = write

    def get(self, key):
        with self.env.begin(write=False) as txn:
            v = txn.get(key.encode())
            return None if v is None else bytes(v)

    def put(self, key, value):
        with self.env.begin(write=True) as txn:
            txn.put(key.encode(), bytes(value))

    def close(self):
        self.env.close()


def open_env(path, write=False):
    if lmdb is not None and os.path.exists(os.path.join(path, "data.mdb")):
        return LmdbEnv(path, write)
    return DirEnv(path)


class PointPaintDataset:
    """PointPaintDataset (point_paint_dataset.py:8-46) over every trajectory under ``data_dir``: ``ds[i]`` ->
    (lidar (n,4) float32, rgbs (ncam,3,H,W) uint8 RGB); ``commit(i, painted)`` writes ``lidar_sem_%05d``."""

    def __init__(self, data_dir, num_cams=3, num_plan=0):
        self.envs, self.index = [], []
        for path in sorted(glob.glob(os.path.join(data_dir, "*"))):
            if not os.path.isdir(path):
                continue
            env = open_env(path, write=True)
            n = env.get("len")
            if n is None:
                continue
            self.envs.append(env)
            for i in range(int(n) - num_plan):                       # basic_dataset.py:77-83 skips the last num_plan frames
                self.index.append((len(self.envs) - 1, i))
        self.num_cams = num_cams

    def __len__(self):
        return len(self.index)

    def __getitem__(self, idx):
        import cv2
        e, i = self.index[idx]
        env = self.envs[e]
        lidar = np.frombuffer(env.get(f"lidar_{i:05d}"), np.float32).reshape(-1, 4)
        rgbs = np.stack([cv2.imdecode(np.frombuffer(env.get(f"rgb_{c}_{i:05d}"), np.uint8), cv2.IMREAD_COLOR) for c in range(self.num_cams)])
        return lidar, np.ascontiguousarray(rgbs[..., ::-1].transpose(0, 3, 1, 2))          # BGR -> RGB, NCHW (:31)

    def commit(self, idx, lidar_painted):
        e, i = self.index[idx]
        self.envs[e].put(f"lidar_sem_{i:05d}", np.ascontiguousarray(lidar_painted).astype(np.float32).tobytes())

    def close(self):
        for e in self.envs:
            e.close()


@torch.no_grad()
def paint_dataset(dataset, seg_model, camera_x=1.5, camera_z=2.4, frames_per_batch=32, device=torch.device("cuda"), progress=None):
    """PointPainter.step (data_paint.py:65-81) for every frame of ``dataset``, batched.  Returns the number of painted frames."""
    seg_model = seg_model.to(device).eval()
    convs = PP.make_converters(camera_x, camera_z)
    cams = np.stack([c.packed() for c in convs])
    done = 0
    for lo in range(0, len(dataset), frames_per_batch):
        items = [dataset[i] for i in range(lo, min(lo + frames_per_batch, len(dataset)))]
        F_ = len(items)
        n_max = max(len(l) for l, _ in items)
        pts = torch.full((F_, max(n_max, 1), 4), float("nan"))
        for f, (l, _) in enumerate(items):
            pts[f, :len(l)] = torch.from_numpy(l)
        rgb = torch.from_numpy(np.concatenate([r for _, r in items]))                       # (F*ncam,3,H,W) uint8
        h, w = rgb.shape[2:]
        feat, table, ncls = seg_model.forward_features_nhwc(rgb.permute(0, 2, 3, 1).contiguous().to(device))
        out = torch.empty((F_, pts.shape[1], ncls - 1), device=device)
        ops.paint_deconv_batched(pts.to(device), feat, ncls, table, cams, 0, out, (h, w))
        out = out.cpu().numpy()
        for f, (l, _) in enumerate(items):
            dataset.commit(lo + f, out[f, :len(l)])
        done += F_
        if progress:
            progress(done, len(dataset))
    return done


def main():
    from .rgb import RGBSegmentationModel
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--seg-weights", required=True, help="state_dict of RGBSegmentationModel (weights/seg_1.th)")
    ap.add_argument("--frames-per-batch", type=int, default=32)
    ap.add_argument("--precision", default="f16", choices=["f16", "fp32"])
    args = ap.parse_args()
    seg = RGBSegmentationModel([4, 6, 7, 10])
    seg.load_state_dict(torch.load(args.seg_weights, map_location="cpu"))
    seg.set_precision(args.precision)
    ds = PointPaintDataset(args.data_dir)
    n = paint_dataset(ds, seg, frames_per_batch=args.frames_per_batch, progress=lambda d, t: print(f"\r{d}/{t}", end="", flush=True))
    ds.close()
    print(f"\npainted {n} frames")


if __name__ == "__main__":
    main()
