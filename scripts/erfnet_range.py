"""one ERFNet forward (3B uint8 frames -> 16-ch decoder map, the product path) inside a cudaProfiler range"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lav_b200 import synth
from tests import util
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
seg, _ = util.seg_model(dev)
seg.set_precision("f16")
rgb = torch.cat([synth.rgb_frames(tag=f"er{b % 4}", smooth=True) for b in range(B)]).to(dev)
with torch.no_grad():
    for _ in range(3):
        seg.forward_features_nhwc(rgb)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    seg.forward_features_nhwc(rgb)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
