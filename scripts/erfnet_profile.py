"""torch.profiler kernel table of one ERFNet forward (3B images, f16)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from lav_b200 import synth
from tests import util
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
seg, _ = util.seg_model(dev)
seg.set_precision("f16")
rgb = synth.rgb_frames().to(dev).repeat(B, 1, 1, 1)
with torch.no_grad():
    for _ in range(3):
        seg.forward_nhwc(rgb)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        seg.forward_nhwc(rgb)
        torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)[:12]:
    print(f"{e.self_device_time_total:10.1f} us  n={e.count:3d}  {e.key[:100]}")
