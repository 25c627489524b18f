"""Round-2 exploration: does processing the conv stacks in L2-resident sub-batches beat whole-batch launches?
  python scripts/r2_explore.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lav_b200 import ops, synth
from tests import util

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")


def graph_time(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


with torch.no_grad():
    seg, _ = util.seg_model(dev)
    seg.set_precision("f16")
    rgb = synth.rgb_frames().to(dev).repeat(B, 1, 1, 1)
    for chunk in (3 * B,):
        if chunk > 3 * B:
            continue
        def run(chunk=chunk):
            for i in range(0, 3 * B, chunk):
                seg.forward_nhwc(rgb[i:i + chunk])
        print(f"ERFNet {3 * B} images in chunks of {chunk}: {graph_time(run):.3f} ms", flush=True)
    lid, _ = util.lidar_model(dev)
    lid.set_precision("f16")
    canvas = (torch.randn(B, 320, 320, 128, device=dev) * 0.3).to(ops.h16())
    for chunk in (B, 16, 8, 4):
        if chunk > B:
            continue
        def run(chunk=chunk):
            for i in range(0, B, chunk):
                f = lid.backbone.forward_nhwc(canvas[i:i + chunk])
                lid.heads_nhwc(f) if hasattr(lid, "heads_nhwc") else None
        print(f"BEV backbone {B} frames in chunks of {chunk}: {graph_time(run):.3f} ms", flush=True)
