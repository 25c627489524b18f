import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lav_b200 import synth
from oracle import lav_ref as O
from tests import util
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
dig = json.load(open("tests/golden/lidar_model_train_grad_digest.json"))
m, sd = util.lidar_model(dev); m.train()
clouds = util.pillar_clouds()
outs = m([c.to(dev) for c in clouds], [len(c) for c in clouds])
gw = [torch.randn(o.shape, generator=synth._gen(7, f"gw{i}")) for i, o in enumerate(outs)]
loss = sum((o * g.to(dev)).sum() for o, g in zip(outs, gw)) / 1e3
loss.backward()
# the same step through the oracle on THIS box's CPU
sd_t = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
o_outs = O.lidar_model(sd_t, clouds, [len(c) for c in clouds], training=True, **util.GRID)
(sum((o * g).sum() for o, g in zip(o_outs, gw)) / 1e3).backward()
print(f"{'param':50s} {'gpu/gold norm':>14s} {'gpu proj err':>13s} {'cpu-here proj err':>17s}")
for k, p in m.named_parameters():
    g = p.grad.detach().cpu(); r = torch.randn(g.shape, generator=synth._gen(13, "dg:" + k))
    n0, s0, p0, mx = dig[k]
    gc = sd_t[k].grad
    print(f"{k:50s} {float(g.norm())/max(n0,1e-12):14.5f} {abs(float((g*r).sum())-p0)/max(n0,1e-12):13.2e} {abs(float((gc*r).sum())-p0)/max(n0,1e-12):17.2e}")
