"""forward + hand-written backward of the rotated crop at the training step's size (32 frames x 384 channels x 160 x 160 fp32,
~60 crops of 96 x 96) inside a cudaProfiler range, with the F.grid_sample path timed beside it by events."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from lav_b200 import ops
from lav_b200.heads import crop_theta
B, C, H, W, S, K = 32, 384, 160, 160, 96, 60
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
feats = torch.randn(B, H, W, C, generator=g).to(dev).requires_grad_(True)
locs = (torch.randn(K, 2, generator=g) * 6).to(dev)
oris = (torch.rand(K, generator=g) * 0.6 - 0.3).to(dev)
fidx = torch.cat([torch.arange(B), torch.randint(0, B, (K - B,), generator=g)]).to(torch.int32).to(dev)
theta = crop_theta(locs, oris, H, W, 2.0, S, torch.tensor(0., device=dev), torch.tensor(0.75, device=dev))
gout = torch.randn(K, S, S, C, generator=g).to(dev)


def ours():
    feats.grad = None
    out = ops.CropBilinear.apply(feats, fidx, theta, S)
    out.backward(gout)


def ref():
    feats.grad = None
    grids = F.affine_grid(theta, torch.Size((K, C, S, S)), align_corners=True)
    out = F.grid_sample(feats.permute(0, 3, 1, 2)[fidx.long()], grids, align_corners=True)
    out.backward(gout.permute(0, 3, 1, 2))


for fn in (ours, ref):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    print(f"{fn.__name__}: forward + backward {a.elapsed_time(b):.2f} ms")
torch.cuda.profiler.start()
ours()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
