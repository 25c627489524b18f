"""one ERFNet forward (B frames x 3 cameras, f16) inside a cudaProfiler range, for an ncu launch list"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lav_b200 import synth
from tests import util
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
seg, _ = util.seg_model(dev)
seg.set_precision("f16")
rgb = synth.rgb_frames().to(dev).repeat(B, 1, 1, 1)
with torch.no_grad():
    for _ in range(2):
        seg.forward_nhwc(rgb)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    seg.forward_nhwc(rgb)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
