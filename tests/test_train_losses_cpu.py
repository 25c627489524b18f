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
