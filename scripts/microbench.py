"""Per-kernel timings on the GPU box (CUDA events, L2 flushed between iterations)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lav_b200 import synth, ops
from lav_b200 import point_painting as PP
from tests import util

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
lidar = synth.lidar_sweep(40000).to(dev)
sem = synth.sem_probs().to(dev)
convs = PP.make_converters()
res["paint_40k_ms"] = timeit(lambda: PP.forward_paint(lidar, sem, convs))
m, _ = util.lidar_model(dev)
base = torch.cat([synth.painted_sweep(40000), torch.tensor([[1., 0, 0]]).expand(40000, 3)], 1).to(dev)
for B in (1, 32):
    batch = base[None].repeat(B, 1, 1).contiguous()
    with torch.no_grad():
        t = timeit(lambda: m.point_pillar_net(batch, [40000] * B))
    res[f"pillar_B{B}_40k_ms"] = t
    res[f"pillar_B{B}_GBs"] = B * (40000 * 44 + 64 * 320 * 320 * 4) / t / 1e6
st = synth.stacked_lidar().to(dev)
with torch.no_grad():
    res["pillar_B1_120k_ms"] = timeit(lambda: m.point_pillar_net([st], [len(st)]))
for prec in ("fp32", "f16"):
    m.set_precision(prec)
    for B in (1, 4):
        canvas = torch.randn(B, 320, 320, 64, device=dev).relu()
        with torch.no_grad():
            tb = timeit(lambda: m.backbone.forward_nhwc(canvas), iters=5)
            feats = m.backbone.forward_nhwc(canvas)
            th = timeit(lambda: m.heads_nhwc(feats), iters=5)
        res[f"backbone_{prec}_B{B}_ms"] = tb
        res[f"backbone_{prec}_B{B}_TFs"] = B * 25.376e9 / tb / 1e9
        res[f"heads_{prec}_B{B}_ms"] = th
        res[f"heads_{prec}_B{B}_TFs"] = B * 45.564e9 / th / 1e9
seg, _ = util.seg_model(dev)
rgb = synth.rgb_frames().to(dev)
for prec in ("fp32", "f16"):
    seg.set_precision(prec)
    with torch.no_grad():
        t = timeit(lambda: seg.forward_nhwc(rgb), iters=5)
    res[f"erfnet_{prec}_3cam_ms"] = t
    res[f"erfnet_{prec}_TFs"] = 22.348e9 / t / 1e9
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)
