"""The eight loss terms of train_lidar against the REFERENCE's own LAV.train_lidar (lav/lav_final_v2.py:177-225 +
lav/models/loss.py:5-27).  tests/golden/train_losses.npz holds what the unmodified reference method returned on the seeded
stand-in model outputs of synth.loss_block_inputs (oracle/pin_against_reference.py, sub-models stubbed)."""
import os

import numpy as np
import pytest
import torch

from lav_b200 import synth
from lav_b200 import train as T


def _check(device, golden_dir, tol):
    gold = np.load(os.path.join(golden_dir, "train_losses.npz"))
    names = [str(n) for n in gold["names"]]
    outs, planner, tg = synth.loss_block_inputs()
    mv = lambda t: t.to(device)
    outs, planner = tuple(map(mv, outs)), tuple(map(mv, planner))
    tg = {k: mv(v) for k, v in tg.items()}
    for distill in (True, False):
        for mode in ("full", "perceive_only", "motion_only"):
            cfg = T.LossConfig(distill=distill, perceive_only=(mode == "perceive_only"), motion_only=(mode == "motion_only"))
            total, parts = T.train_losses(outs, planner, tg["heatmaps"], tg["sizemaps"], tg["orimaps"], tg["bev"], tg["ego_locs"], tg["cmds"],
                                          tg["bras"], T.build_seg_mask().to(device), cfg)
            want = gold[f"{'distill' if distill else 'nodistill'}_{mode}"]
            got = np.array([float(parts[n]) for n in names])
            np.testing.assert_allclose(got, want, rtol=tol, atol=tol * 1e-2)
            w = dict(zip(names, want))
            det = w["hm_loss"] + cfg.box_weight * w["box_loss"] + cfg.ori_weight * w["ori_loss"] + w["seg_loss"]
            mot = w["plan_loss"] + w["ego_cast_loss"] + cfg.other_weight * w["other_cast_loss"] + cfg.cmd_weight * w["cmd_loss"]
            tot = det if cfg.perceive_only else mot if cfg.motion_only else mot + cfg.perception_weight * det      # lav_final_v2.py:215-220
            assert abs(float(total) - tot) <= tol * abs(tot) * 4


def test_train_losses_match_reference_golden(golden_dir):
    _check(torch.device("cpu"), golden_dir, 1e-6)


@pytest.mark.gpu
def test_train_losses_match_reference_golden_on_gpu(cuda, golden_dir):
    _check(cuda, golden_dir, 2e-5)


def test_detloss_module_signature():
    g = torch.Generator().manual_seed(0)
    ph, hm = torch.randn(2, 2, 16, 16, generator=g), torch.rand(2, 2, 16, 16, generator=g)
    a = T.DetLoss()(ph, hm, ph, hm, ph, hm)
    b = T.detection_losses(ph, hm, ph, hm, ph, hm)
    assert all(torch.equal(x, y) for x, y in zip(a, b)) and len(a) == 3


def test_fused_heads_trunk_equals_per_head_layers():
    """train.py:_heads_trunk_fused (one 4x-wide conv + one batch norm) against the heads' own Conv -> ReLU -> BatchNorm:
    outputs, gradients (features, conv and BN parameters) and the running statistics written back."""
    import copy
    from lav_b200.lidar import Head
    from lav_b200 import train as T
    torch.manual_seed(3)
    heads = [Head(24, n, num_hidden=8) for n in (1, 2, 2, 3)]
    for h in heads:
        h.train()
        h.net[2].weight.data.uniform_(0.5, 1.5); h.net[2].bias.data.normal_()
    ref = copy.deepcopy(heads)
    feats = torch.randn(3, 24, 10, 12)
    fa, fb = feats.clone().requires_grad_(True), feats.clone().requires_grad_(True)
    hidden = T._heads_trunk_fused(heads, fa)
    assert hidden is not None
    got = [h.net[3](x) for h, x in zip(heads, hidden)]
    want = [h.net(fb) for h in ref]
    sum(o.square().sum() for o in got).backward()
    sum(o.square().sum() for o in want).backward()
    for g, w in zip(got, want):
        assert torch.allclose(g, w, rtol=1e-4, atol=1e-5)
    assert torch.allclose(fa.grad, fb.grad, rtol=1e-3, atol=1e-4)
    for h, r in zip(heads, ref):
        for (n, p), (_, q) in zip(h.named_parameters(), r.named_parameters()):
            assert torch.allclose(p.grad, q.grad, rtol=1e-3, atol=1e-4), n
        assert torch.allclose(h.net[2].running_mean, r.net[2].running_mean, atol=1e-6)
        assert torch.allclose(h.net[2].running_var, r.net[2].running_var, atol=1e-6)
        assert int(h.net[2].num_batches_tracked) == int(r.net[2].num_batches_tracked) == 1
    heads[1].net[2].eps = 1e-2                                 # heads that differ fall back to the per-head path
    assert T._heads_trunk_fused(heads, fa) is None
