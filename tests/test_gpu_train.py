"""GPU parity of the training path: train-mode LiDARModel forward/backward (CUDA pillar decorate + scatter-max with
arg-routed backward, batch-stat BatchNorm) against the REFERENCE loss/gradients stored in tests/golden."""
import os

import numpy as np
import pytest
import torch

from lav_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu


def test_train_forward_backward_matches_reference_golden(cuda, golden_dir):
    gold = np.load(os.path.join(golden_dir, "lidar_model_train.npz"))
    m, _ = util.lidar_model(cuda)
    m.train()
    clouds = [c.to(cuda) for c in util.pillar_clouds()]
    outs = m(clouds, [len(c) for c in clouds])
    names = ["features", "center", "box", "ori", "seg"]
    for n, o in zip(names, outs):
        assert util.rel_err(o.detach()[:, :, ::8, ::8], torch.from_numpy(gold["out_" + n])) < 1e-3, n
    gw = [torch.randn(o.shape, generator=synth._gen(7, f"gw{i}")).to(cuda) for i, o in enumerate(outs)]
    loss = sum((o * g).sum() for o, g in zip(outs, gw)) / 1e3
    assert abs(float(loss) - float(gold["loss"])) < 1e-3 * abs(float(gold["loss"])) + 1e-3
    loss.backward()
    grads = dict(m.named_parameters())
    for key in gold.files:
        if key.startswith("grad:"):
            g = grads[key[5:]].grad
            assert g is not None, key
            # gradients that travelled the whole backward chain (17 conv layers, batch-stat BN, arg-max routing that can
            # flip on near ties between cuDNN-GPU and CPU roundings) get a wider band than the shallow ones
            tol = 2e-2 if "point_pillar_net" in key else 5e-3
            assert util.rel_err(g, torch.from_numpy(gold[key])) < tol, (key, util.rel_err(g, torch.from_numpy(gold[key])))


def test_perception_trainer_step_decreases_loss(cuda):
    from lav_b200.train import PerceptionTrainer
    m, _ = util.lidar_model(cuda)
    tr = PerceptionTrainer(m, lr=1e-3, device=cuda)
    clouds = [c.to(cuda) for c in util.pillar_clouds()]
    g = synth._gen(3, "tgt")
    heat = (torch.rand(2, 2, 320, 320, generator=g) > 0.995).float().to(cuda)
    size = torch.rand(2, 2, 320, 320, generator=g).to(cuda)
    ori = torch.randn(2, 2, 320, 320, generator=g).to(cuda)
    bev = (torch.rand(2, 9, 320, 320, generator=g) > 0.5).float().to(cuda)
    losses = [float(tr.train_step(clouds, [len(c) for c in clouds], heat, size, ori, bev)[0]) for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
