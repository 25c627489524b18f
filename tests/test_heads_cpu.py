"""CPU: the PyTorch heads (UniPlanner / brake) reproduce the REFERENCE outputs in tests/golden and the oracle,
and the host-side detection decode keeps the reference's filters."""
import os

import numpy as np
import torch

from lav_b200 import synth
from lav_b200.heads import BEVPlanner, RGBBrakePredictionModel, UniPlanner
from oracle import lav_ref as O
from tests import util

KW = dict(pixels_per_meter=4, crop_size=96, feature_x_jitter=1.5, feature_angle_jitter=20, x_offset=0,
          y_offset=1 + (-10) / ((70 + 10) / 2), num_cmds=6, num_plan=20, num_plan_iter=5)


def uniplanner():
    up = UniPlanner(BEVPlanner(num_frame_stack=2, **KW), num_input_feature=384, **KW).eval()
    sd = synth.fill_state_dict_(up.state_dict())
    up.load_state_dict(sd)
    return up, {k: v.clone() for k, v in sd.items()}


def test_uniplanner_infer_matches_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "uniplanner.npz"))
    up, sd = uniplanner()
    _, lsd = util.lidar_model()
    clouds = util.pillar_clouds()
    with torch.no_grad():
        feats = O.lidar_model(lsd, clouds, [len(c) for c in clouds], **util.GRID)[0][0]
        det = [tuple(r) for r in gold["det"]]
        epl, ecl, ocl, occ = up.infer(feats, det, 2, torch.tensor([0.0, -20.0]))
    sc = float(np.abs(gold["other_cast"]).max()) + 1
    np.testing.assert_allclose(epl.numpy(), gold["ego_plan"], atol=2e-4 * sc)
    np.testing.assert_allclose(ecl.numpy(), gold["ego_cast"], atol=2e-4 * sc)
    np.testing.assert_allclose(ocl.numpy(), gold["other_cast"], atol=2e-4 * sc)
    np.testing.assert_allclose(occ.numpy(), gold["other_cmds"], atol=1e-5)
    # zero detections: reference returns empty CPU tensors
    epl0, _, ocl0, occ0 = up.infer(feats, [], 2, torch.tensor([0.0, -20.0]))
    assert ocl0.shape == (0, 6, 20, 2) and occ0.shape == (0, 6)
    np.testing.assert_allclose(epl0.numpy(), gold["ego_plan"], atol=2e-4 * sc)


def test_brake_matches_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "brake.npz"))
    m = RGBBrakePredictionModel([4, 6, 7, 10]).eval()
    m.load_state_dict(synth.fill_state_dict_(m.state_dict()))
    rgb1 = synth.rgb_frames(smooth=True, tag="wide", n_cam=1, h=288, w=768).permute(0, 3, 1, 2).float()
    rgb2 = synth.rgb_frames(smooth=True, tag="tele", n_cam=1, h=192, w=480).permute(0, 3, 1, 2).float()
    with torch.no_grad():
        got = m(rgb1, rgb2)
    np.testing.assert_allclose(got.numpy(), gold["seeded"], atol=1e-5)


def test_det_decode_matches_reference_golden(golden_dir):
    from lav_b200.model_inference import InferModel
    gold = np.load(os.path.join(golden_dir, "uniplanner.npz"))
    gb = synth._gen(11, "blobs")
    yy, xx = torch.meshgrid(torch.arange(320.), torch.arange(320.), indexing="ij")
    heat = torch.full((2, 320, 320), -6.0)
    centres = [(100, 200), (161, 281), (250, 150), (30, 30), (160, 100), (200, 260), (120, 250)]
    for ci, (cx_, cy_) in enumerate(centres):
        amp = 4.0 + float(torch.rand(1, generator=gb)) * 6
        heat[ci % 2] = torch.maximum(heat[ci % 2], -6 + amp * torch.exp(-((xx - cx_) ** 2 + (yy - cy_) ** 2) / 8.0))
    sizem = torch.rand(2, 320, 320, generator=gb) * 3
    orim = torch.randn(2, 320, 320, generator=gb)
    stub = type("S", (), {"pixels_per_meter": 4})()
    dets = InferModel.decode_packed(stub, InferModel.pack_peaks(torch.sigmoid(heat)[None], sizem[None], orim[None]))[0]
    want = O.det_inference(torch.sigmoid(heat), sizem, orim)
    assert dets == want
    np.testing.assert_allclose(np.array(dets[0]).reshape(-1, 6), gold["det0"], rtol=0, atol=0)
    np.testing.assert_allclose(np.array(dets[1]).reshape(-1, 6), gold["det1"], rtol=0, atol=0)


def test_uniplanner_train_forward_matches_reference_golden(golden_dir):
    """student + frozen teacher training forward (lav/models/uniplanner.py:56-151), seeded jitter, vs the reference run."""
    gold = np.load(os.path.join(golden_dir, "uniplanner_train.npz"))
    up, _ = uniplanner()
    up.train()
    _, lsd = util.lidar_model()
    clouds = util.pillar_clouds()
    with torch.no_grad():
        feats = O.lidar_model(lsd, clouds, [len(c) for c in clouds], **util.GRID)[0][:2] * 0.5
    gt = synth._gen(23, "uptrain")
    bev = (torch.rand(2, 9, 320, 320, generator=gt) > 0.7).float()
    ego_locs = torch.cumsum(torch.rand(2, 21, 2, generator=gt) * torch.tensor([0.2, -1.0]), dim=1)
    locs = torch.randn(2, 6, 21, 2, generator=gt) * 6 + torch.tensor([0.0, -8.0])
    locs[:, 0] = ego_locs
    oris = torch.rand(2, 6, generator=gt) * 0.6 - 0.3
    typs = torch.tensor([[1, 1, 1, 0, 1, 1], [1, 1, 0, 1, 1, 0]])
    nxps = torch.tensor([[0.0, -20.0], [3.0, -15.0]])
    torch.manual_seed(1234)
    out = up(feats, bev, ego_locs, locs, oris, nxps, typs)
    names = ["other_locs", "other_cast_locs", "other_cast_cmds", "other_cast_locs_expert", "other_cast_cmds_expert", "ego_locs",
             "ego_plan_locs", "ego_cast_locs", "ego_cast_cmds", "ego_cast_locs_expert", "ego_plan_locs_expert"]
    assert len(out) == 11
    for n, o in zip(names, out):
        np.testing.assert_allclose(o.detach().numpy(), gold[n], atol=5e-4 * (np.abs(gold[n]).max() + 1), err_msg=n)
    assert out[6].requires_grad and not out[9].requires_grad          # student carries grad, teacher does not


def test_cast_module_path_equals_oracle_cast():
    """heads._cast_branches on CPU tensors (the module path that training and non-CUDA callers take; the CUDA kernel is compared
    with it in tests/test_gpu_frame.py) == the oracle's cast of uniplanner.py:286-301, for both planners."""
    up, sd = uniplanner()
    embd = torch.randn(5, 512, generator=torch.Generator().manual_seed(3)) * 0.6
    with torch.no_grad():
        got = up.cast(embd)
        want = O.up_cast(sd, embd)
        assert got.shape == (5, 6, 20, 2)
        assert util.rel_err(got, want) < 1e-5
        t = up.bev_planner.cast(embd)
        branches = [torch.cumsum(m(g(embd.expand(up.bev_planner.num_plan, 5, -1).permute(1, 0, 2).contiguous())[0]), dim=1)
                    for g, m in zip(up.bev_planner.cast_grus, up.bev_planner.cast_mlps)]
        assert torch.allclose(t, torch.stack(branches, dim=1), atol=1e-6)
