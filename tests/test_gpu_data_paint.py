"""f3 — the offline painter (lav/data_paint.py:44-107) on the reference's on-disk key layout
(lav/utils/datasets/basic_dataset.py:52-53,82-101; point_paint_dataset.py:34-46)."""
import os

import numpy as np
import pytest
import torch

from lav_b200 import synth
from oracle import lav_ref as O
from tests import util

pytestmark = pytest.mark.gpu


def _make_store(root, frames):
    cv2 = pytest.importorskip("cv2")
    from lav_b200.data_paint import DirEnv
    truth = []
    for ti, n_frames in enumerate(frames):
        env = DirEnv(os.path.join(root, f"traj{ti:02d}"))
        env.put("len", str(n_frames).encode())
        env.put("town", b"Town01")
        for i in range(n_frames):
            lidar = synth.lidar_sweep(3000 + 517 * i + 1000 * ti, tag=f"dp{ti}{i}").numpy()
            rgb = synth.rgb_frames(tag=f"dp{ti}{i}", smooth=True).numpy()                   # (3,288,256,3) RGB
            env.put(f"lidar_{i:05d}", lidar.astype(np.float32).tobytes())
            for c in range(3):
                ok, buf = cv2.imencode(".png", np.ascontiguousarray(rgb[c][..., ::-1]))      # stored BGR, lossless
                assert ok
                env.put(f"rgb_{c}_{i:05d}", buf.tobytes())
            truth.append((lidar, rgb))
    return truth


def test_offline_painter_writes_reference_format(cuda, tmp_path):
    from lav_b200.data_paint import DirEnv, PointPaintDataset, paint_dataset
    truth = _make_store(str(tmp_path), [3, 2])
    m, sd = util.seg_model(cuda)
    ds = PointPaintDataset(str(tmp_path))
    assert len(ds) == 5
    lidar0, rgbs0 = ds[0]
    assert lidar0.shape == truth[0][0].shape and rgbs0.shape == (3, 3, 288, 256)
    assert np.array_equal(rgbs0.transpose(0, 2, 3, 1), truth[0][1])                        # PNG round trip, BGR -> RGB
    assert paint_dataset(ds, m, frames_per_batch=3, device=cuda) == 5
    ds.close()
    convs = O.default_converters()
    k = 0
    for ti, n_frames in enumerate([3, 2]):
        env = DirEnv(os.path.join(str(tmp_path), f"traj{ti:02d}"))
        for i in range(n_frames):
            lidar, rgb = truth[k]
            k += 1
            got = np.frombuffer(env.get(f"lidar_sem_{i:05d}"), np.float32).reshape(-1, 4)
            assert got.shape == (len(lidar), 4)
            with torch.no_grad():      # PointPainter.step, data_paint.py:70-78 (fp64 numpy painter)
                sems = torch.softmax(O.erfnet(sd, torch.from_numpy(rgb).permute(0, 3, 1, 2).float()), 1)
            want = O.point_painting_f64(lidar, O.suppress_background(sems).numpy(), convs)
            bad = np.abs(got - want).max(1) > 2e-3
            assert bad.sum() <= max(2, len(lidar) // 1000), f"{bad.sum()} painted rows differ"  # fp32-vs-fp64 pixel-boundary flips
            assert np.array_equal((got != 0).any(1)[~bad], (want != 0).any(1)[~bad])
