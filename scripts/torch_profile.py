"""Kernel-level time table (torch.profiler / CUPTI, eager launches) of the PyTorch-side stages: brake model and planner graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from lav_b200 import synth
from lav_b200.agent import StaticFramePipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
(seg, lid, uni, bra), _ = bench.build_models()
pipe = StaticFramePipeline(seg, lid, uni, bra, B, synth.SWEEP_POINTS, device=dev, precision="f16", use_graphs=False)
rgbs, tels, lidars, prev, poses = bench.synth_frames(B)
nxps = torch.tensor([[0.0, -20.0]] * B); cmds = torch.tensor([3] * B)
for _ in range(2):
    pipe.step(rgbs.to(dev), tels.to(dev), torch.stack(lidars).to(dev), nxps, cmds, fixed_dets=bench.FIXED_DETS)
torch.cuda.synchronize()
K = 3 * B
g, o, s2 = pipe._g2[K]


def table(name, fn):
    with torch.no_grad():
        fn(); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn(); torch.cuda.synchronize()
    print(f"==== {name} (B={B})")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90))


table("brake", lambda: pipe._brake())
table("planner graph body", lambda: pipe._g2_body(K, s2["locs"], s2["oris"], s2["fidx"]))
table("erfnet (3B images)", lambda: pipe.seg_model.forward_nhwc(pipe.rgbs.view(B * 3, 288, 256, 3)))
table("pillars + backbone + heads", lambda: pipe.infer_model.lidar_model.forward_nhwc(pipe.stacked, [pipe.T * pipe.N] * B))
