"""one launch of the fused 4-head conv (384 -> 256, 3x3, 160x160) on B frames inside a cudaProfiler range"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lav_b200 import ops
from lav_b200.layers import TapConv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
wgt = torch.randn(256, 384, 3, 3, device=dev) * 0.02
layer = TapConv(wgt, False, 1, 1, 1, 0, None, pre_relu=True, scale=torch.ones(256, device=dev), shift=torch.zeros(256, device=dev))
x = torch.randn(B, 160, 160, 384, device=dev).to(ops.h16())
out = torch.empty(B, 160, 160, 256, device=dev, dtype=ops.h16())
for _ in range(3):
    layer(x, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
layer(x, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
