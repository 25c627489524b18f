"""Print error metrics of the f16 path (and fp32 path) against the oracle for each LiDARModel output."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lav_ref as O
from tests import util

dev = torch.device("cuda:0")
m, sd = util.lidar_model(dev)
clouds = util.pillar_clouds()
npts = [len(c) for c in clouds]
with torch.no_grad():
    want = O.lidar_model(sd, clouds, npts, **util.GRID)
    want = list(want[:4]) + [torch.logit(want[4].clamp(1e-7, 1 - 1e-7))]
    for prec in ("fp32", "f16"):
        m.set_precision(prec)
        got = m([c.to(dev) for c in clouds], npts)
        got = [g.float().cpu() for g in got]
        got[4] = torch.logit(got[4].clamp(1e-7, 1 - 1e-7))
        for n, a, b in zip(["features", "center", "box", "ori", "seg_logit"], got, want):
            e = (a - b)
            rms = float((e ** 2).mean().sqrt() / (b ** 2).mean().sqrt())
            mx = float(e.abs().max() / b.abs().max())
            band = float(((e.abs() <= 1e-2 * b.abs() + 1e-2 * (b ** 2).mean().sqrt())).float().mean())
            print(f"{prec} {n:10s} rms_rel={rms:.3e} max_err/max_ref={mx:.3e} frac_within(1e-2*|ref|+1e-2*rms)={band:.5f} ref_rms={float((b**2).mean().sqrt()):.3f}")
