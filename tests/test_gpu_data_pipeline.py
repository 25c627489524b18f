"""f4 (device-side part): the LiDAR stacking of TemporalLiDARPaintedDataset.__getitem__
(lav/utils/datasets/temporal_lidar_painted_dataset.py:24-93) and detections_to_heatmap (lidar_dataset.py:92-127) against
numpy restatements of those lines."""
import math

import numpy as np
import pytest
import torch

from lav_b200 import synth
from oracle import lav_ref as O

pytestmark = pytest.mark.gpu


def _ref_stack(sweeps, angle, jitters, convs):
    """numpy restatement of temporal_lidar_painted_dataset.py:35-87 + lidar_dataset.py:14-23,175-182 (no shuffle)."""
    loc0, ori0 = sweeps[0][2], sweeps[0][3]
    out = []
    for i, (xyzr, painted, loc, ori) in enumerate(sweeps):
        keep = ~((xyzr[:, 0] > -2.4) & (xyzr[:, 0] < 0) & (xyzr[:, 1] > -0.8) & (xyzr[:, 1] < 0.8) & (xyzr[:, 2] > -1.5) & (xyzr[:, 2] < -1))
        xyzr, painted = xyzr[keep], painted[keep].copy()
        rad = np.deg2rad(-angle)
        xyzr = xyzr @ np.array([[np.cos(rad), np.sin(rad), 0, 0], [-np.sin(rad), np.cos(rad), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        mask = O.point_painting_f64(xyzr, np.ones((3, 1, 288, 256)), convs)
        painted *= mask
        lj, oj = (np.zeros(2), 0.0) if i == 0 else jitters[i]
        dloc = (loc - loc0 + lj) @ np.array([[np.cos(ori0), -np.sin(ori0)], [np.sin(ori0), np.cos(ori0)]])
        d = ori + oj - ori0
        xyzr = xyzr @ np.array([[np.cos(d), np.sin(d), 0, 0], [-np.sin(d), np.cos(d), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        xyzr[:, :2] += dloc
        t = np.zeros((len(xyzr), len(sweeps)))
        t[:, i] = 1
        out.append(np.concatenate([xyzr, painted, t], 1))
    return np.concatenate(out)


def test_gpu_lidar_stacker_matches_dataset_lines(cuda):
    from lav_b200.data_pipeline import GpuLidarStacker
    loc, ori = synth.ego_motion(3, tag="dl")
    sweeps = []
    for i in range(3):
        s = synth.painted_sweep(9000 + 700 * i, tag=f"dl{i}").numpy().astype(np.float64)
        s[:300, :3] = np.array([-1.0, 0.1, -1.2]) + 0.3 * np.random.RandomState(i).randn(300, 3)      # some roof points
        s[:, 3] = (np.arange(len(s)) + 0.5) / len(s)                                                  # unique "intensity" = row id (the stacker shuffles)
        sweeps.append((s[:, :4].astype(np.float32), s[:, 4:].astype(np.float32), np.asarray(loc[i], dtype=np.float64), float(ori[i])))
    jit = [None, (np.array([0.05, -0.02]), 0.01), (np.array([-0.03, 0.04]), -0.02)]
    angle = 7.5
    st = GpuLidarStacker(max_lidar_points=40000, device=cuda)
    got, num = st(sweeps, angle, jit, generator=torch.Generator().manual_seed(0))
    want = _ref_stack([(a.astype(np.float64), b.astype(np.float64), c, d) for a, b, c, d in sweeps], angle, jit, O.default_converters())
    assert num == len(want) <= 40000
    g = got[:num].cpu().numpy()
    assert np.all(g.sum(0)[8:] == want.sum(0)[8:])                                      # one-hot counts per sweep
    # the stacker shuffles: compare as sets of rows (sort both by the raw intensity column, unique per synthetic point, + time index)
    key = lambda a: np.lexsort((a[:, 3], a[:, 8], a[:, 9], a[:, 10]))
    g, w = g[key(g)], want[key(want)]
    assert np.abs(g[:, :4] - w[:, :4]).max() < 2e-4                                     # fp32 rotation chain vs fp64 numpy
    flips = (np.abs(g[:, 4:8] - w[:, 4:8]).max(1) > 1e-6).sum()
    assert flips <= max(2, num // 2000), flips                                          # FOV-boundary flips of the fp32 projection
    assert float(got[num:].abs().max()) == 0.0 if num < 40000 else True


def test_detections_to_heatmap_matches_dataset_lines(cuda):
    from lav_b200.data_pipeline import detections_to_heatmap
    rs = np.random.RandomState(3)
    locs = np.stack([rs.uniform(-30, 30, 9), rs.uniform(-60, 5, 9)], 1)
    oris = rs.uniform(-3, 3, 9)
    bbox = rs.uniform(0.5, 3, (9, 2))
    typs = np.array([1, 1, 0, 1, 0, 1, 1, 0, 1])
    h = w = 320
    heat, size, orim = torch.zeros((2, h, w)), torch.zeros((2, h, w)), torch.zeros((2, h, w))
    for i in (0, 1):                                      # lidar_dataset.py:98-125
        idx = typs == i
        loc, ori, box = (torch.tensor(a[idx], dtype=torch.float32) for a in (locs, oris, bbox))
        x, y = torch.arange(w), torch.arange(h)
        cx, cy = -loc[:, 0] * 4 + 80 * 4 / 2, -loc[:, 1] * 4 + h + (-10) * 4
        gx = (-((x[:, None] - cx[None, :]) / 1) ** 2).exp()
        gy = (-((y[:, None] - cy[None, :]) / 1) ** 2).exp()
        gaussian, who = (gx[None] * gy[:, None]).max(dim=-1)
        mask = gaussian > heat.max(dim=0)[0]
        size[:, mask] = box.T[:, who[mask]] * 4
        orim[0, mask] = torch.cos(ori[who[mask]])
        orim[1, mask] = torch.sin(ori[who[mask]])
        heat[i] = gaussian
    g_heat, g_size, g_ori = detections_to_heatmap(locs, oris, bbox, typs, device=cuda)
    assert torch.allclose(g_heat.cpu(), heat, atol=1e-6)
    assert torch.allclose(g_size.cpu(), size, atol=1e-5) and torch.allclose(g_ori.cpu(), orim, atol=1e-5)
