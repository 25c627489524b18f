"""Device-side pieces of the training data pipeline — the per-sample LiDAR work and the detection-target rasterisation of
`TemporalLiDARPaintedDataset.__getitem__` (lav/utils/datasets/temporal_lidar_painted_dataset.py:13-179) and
`LiDARDataset.detections_to_heatmap` (lav/utils/datasets/lidar_dataset.py:92-127).

In the reference these run in 16 numpy / OpenCV DataLoader workers per GPU; at 8 x B200 the loader, not the GPUs, bounds the
training step.  Here the record reads stay on the host (key-value store, see data_paint.py) and everything that
touches the points runs on the GPU with the frame path's own kernels:

    roof filter (order-preserving)         LiDARDataset.preprocess, lidar_dataset.py:14-23          lavb_roof_filter
    rotation jitter                        rotate_lidar(-angle), :175-182                           lavb_stack_sweep (R only)
    re-mask painted features by camera FOV point_painting(xyzr, ones) after the rotation, :58-60     lavb_paint (mode 0 on a ones map)
    ego-motion + jitter + time one-hot     move_lidar_points, :159-177 / :62-66,80-87                lavb_stack_sweep
    shuffle, truncate to max_lidar_points, zero-pad  :89-91,131-133                                  torch.randperm + copies
    heat / size / orientation maps         detections_to_heatmap                                    vectorised torch on the device
Not here (host, as in the reference): LMDB reads, the actor filter, BEV image loading and its cv2 rotation.
"""
import math

import numpy as np
import torch

from . import ops
from . import point_painting as PP


class GpuLidarStacker:
    """The LiDAR half of TemporalLiDARPaintedDataset.__getitem__ for one sample, on ``device``."""

    def __init__(self, num_frame_stack=2, seg_channels=4, max_lidar_points=120000, camera_x=1.5, camera_z=2.4, rgb_hw=(288, 256),
                 device=torch.device("cuda")):
        self.T, self.C, self.max_points, self.device = num_frame_stack + 1, seg_channels, max_lidar_points, device
        self.cams = np.stack([c.packed() for c in PP.make_converters(camera_x, camera_z, rgb_hw[0], rgb_hw[1])])
        self.ones = torch.ones((len(self.cams), 1, rgb_hw[0], rgb_hw[1]), device=device)       # `self.dummy` of the reference

    @torch.no_grad()
    def __call__(self, sweeps, angle_deg, jitters=None, generator=None):
        """sweeps: [(xyzr (n,4), painted (n,C), ego_loc (2,), ego_ori)] newest first (index, index-1, ...), numpy or tensors;
        angle_deg: the sample's rotation jitter; jitters[i] = (loc_jitter (2,), ori_jitter) of sweep i (zeros for i = 0).
        Returns (lidar (max_points, 4+C+T) fp32 zero-padded, num_points)."""
        dev = self.device
        loc0, ori0 = np.asarray(sweeps[0][2], dtype=np.float64), float(sweeps[0][3])
        rad = math.radians(-angle_deg)
        R_aug = np.array([[math.cos(rad), math.sin(rad), 0], [-math.sin(rad), math.cos(rad), 0], [0, 0, 1]], dtype=np.float32)
        rows = []
        for i, (xyzr, painted, loc, ori) in enumerate(sweeps):
            raw = torch.cat([torch.as_tensor(xyzr, dtype=torch.float32), torch.as_tensor(painted, dtype=torch.float32)], 1).to(dev).contiguous()
            kept, cnt = ops.roof_filter(raw)                                                 # preprocess (both arrays, same rows)
            kept = kept[:int(cnt[0])].contiguous()
            n = kept.shape[0]
            rot = torch.empty((n, 4 + self.C), device=dev)
            ops.stack_sweep(kept, R_aug, 0.0, 0.0, 0, 0, rot)                                # rotate_lidar(xyzr, -angle)
            vis = ops.paint(rot, self.ones, self.cams, mode=0)                               # 1 where some camera still sees the point
            rot[:, 4:] *= vis
            lj, oj = (np.zeros(2), 0.0) if (jitters is None or i == 0) else jitters[i]
            dloc = (np.asarray(loc, dtype=np.float64) - loc0 + np.asarray(lj)) @ np.array([[math.cos(ori0), -math.sin(ori0)], [math.sin(ori0), math.cos(ori0)]])
            d = float(ori) + float(oj) - ori0
            R_mv = np.array([[math.cos(d), math.sin(d), 0], [-math.sin(d), math.cos(d), 0], [0, 0, 1]], dtype=np.float32)
            out = torch.empty((n, 4 + self.C + self.T), device=dev)
            ops.stack_sweep(rot, R_mv, dloc[0], dloc[1], i, self.T, out)                     # move_lidar_points + one-hot(t)
            rows.append(out)
        lidar = torch.cat(rows)
        total = lidar.shape[0]
        perm = torch.randperm(total, generator=generator, device=dev if generator is None or generator.device.type == "cuda" else "cpu").to(dev)
        lidar = lidar[perm[:self.max_points]]
        padded = torch.zeros((self.max_points, lidar.shape[1]), device=dev)
        num = min(self.max_points, total)
        padded[:num] = lidar[:num]
        return padded, num


@torch.no_grad()
def detections_to_heatmap(locs, oris, bbox, typs, min_x=-10, max_x=70, min_y=-40, max_y=40, pixels_per_meter=4, radius=1,
                          device=torch.device("cuda")):
    """LiDARDataset.detections_to_heatmap (lidar_dataset.py:92-127) on the device.  locs (N,2) ego-frame metres, oris (N,),
    bbox (N,2), typs (N,) in {0: pedestrian, 1: vehicle} -> heatmap (2,h,w), sizemap (2,h,w), orimap (2,h,w)."""
    h, w = (max_y - min_y) * pixels_per_meter, (max_x - min_x) * pixels_per_meter              # len(y_edges), len(x_edges)
    heat = torch.zeros((2, h, w), device=device)
    size = torch.zeros((2, h, w), device=device)
    orim = torch.zeros((2, h, w), device=device)
    locs, oris, bbox = (torch.as_tensor(np.asarray(t), dtype=torch.float32, device=device) for t in (locs, oris, bbox))
    typs = torch.as_tensor(np.asarray(typs), device=device)
    x = torch.arange(w, device=device, dtype=torch.float32)
    y = torch.arange(h, device=device, dtype=torch.float32)
    for i in (0, 1):
        sel = typs == i
        if int(sel.sum()) == 0:
            continue
        loc, ori, box = locs[sel], oris[sel], bbox[sel]
        cx = -loc[:, 0] * pixels_per_meter + (max_y - min_y) * pixels_per_meter / 2
        cy = -loc[:, 1] * pixels_per_meter + h + min_x * pixels_per_meter
        gx = torch.exp(-((x[:, None] - cx[None, :]) / radius) ** 2)                              # (w, K)
        gy = torch.exp(-((y[:, None] - cy[None, :]) / radius) ** 2)                              # (h, K)
        gaussian, who = (gx[None] * gy[:, None]).max(dim=-1)                                     # (h, w): best actor per pixel
        mask = gaussian > heat.max(dim=0)[0]
        size[:, mask] = box.T[:, who[mask]] * pixels_per_meter
        orim[0, mask] = torch.cos(ori[who[mask]])
        orim[1, mask] = torch.sin(ori[who[mask]])
        heat[i] = gaussian
    return heat, size, orim
