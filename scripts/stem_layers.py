"""brake-model stem + pool on raw camera bytes (B frames) inside a cudaProfiler range, for ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lav_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
(seg, lid, uni, bra), _ = bench.build_models()
bra = bra.to(dev).eval()
bra.conv_backbone.to(ops.h16()).to(memory_format=torch.channels_last)
bra.attn1.to(ops.h16()); bra.attn2.to(ops.h16())
g = torch.Generator().manual_seed(1)
rgbs = torch.randint(0, 256, (B, 3, 288, 256, 3), generator=g, dtype=torch.uint8).to(dev)
tel = torch.randint(0, 256, (B, 192, 480, 3), generator=g, dtype=torch.uint8).to(dev)
with torch.no_grad():
    for _ in range(2):
        bra.forward_u8(rgbs, tel)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    bra.forward_u8(rgbs, tel)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
