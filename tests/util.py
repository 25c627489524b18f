"""Shared builders for the tests: seeded weights (identical in the pin script) and inputs."""
import os

import numpy as np
import torch

from lav_b200 import synth

GRID = dict(min_x=-10, max_x=70, min_y=-40, max_y=40)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lidar_model(device=None):
    from lav_b200.lidar import LiDARModel
    m = LiDARModel(num_input=16, num_features=[64, 64], backbone="cnn", pixels_per_meter=4, **GRID).eval()
    sd = synth.fill_state_dict_(m.state_dict())
    m.load_state_dict(sd)
    if device is not None:
        m = m.to(device)
    return m, {k: v.clone() for k, v in sd.items()}


def seg_model(device=None, real=False):
    from lav_b200.rgb import RGBSegmentationModel
    m = RGBSegmentationModel([4, 6, 7, 10]).eval()
    if real:
        sd = torch.load(os.path.join(ROOT, "oracle", "_ref", "seg_1.state_dict.pt"), map_location="cpu")
    else:
        sd = synth.fill_state_dict_(m.state_dict())
    m.load_state_dict(sd)
    if device is not None:
        m = m.to(device)
    return m, {k: v.clone() for k, v in sd.items()}


def have_real_seg():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "seg_1.state_dict.pt"))


def pillar_clouds():
    return [synth.stacked_lidar(2000, tag="pp0"), synth.stacked_lidar(1500, tag="pp1")]


def paint_inputs():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "paint.npz"))
    lidar = torch.cat([synth.lidar_sweep(8192, tag="paint"), torch.from_numpy(gold["edge"])]).contiguous()
    return lidar, synth.sem_probs(tag="paint"), gold


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))
