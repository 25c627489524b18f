"""Per-stage GPU time of one tick (each stage captured as its own CUDA graph and replayed), B agents."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lav_b200 import ops, synth
from lav_b200.agent import StaticFramePipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
(seg, lid, uni, bra), _ = bench.build_models()
pipe = StaticFramePipeline(seg, lid, uni, bra, B, synth.SWEEP_POINTS, device=dev, precision="f16", use_graphs=True)
rgbs, tels, lidars, prev, poses = bench.synth_frames(B)
pipe.tick = 10
for b in range(B):
    loc, ori = poses[b]
    pipe.preload_history(b, [(prev[b][k % 2].to(dev), loc[1 + (k % 2)], ori[1 + (k % 2)]) for k in range(10)])
nxps = torch.tensor([[0.0, -20.0]] * B); cmds = torch.tensor([3] * B)
out = pipe.step(rgbs.to(dev), tels.to(dev), torch.stack(lidars).to(dev), nxps, cmds, fixed_dets=bench.FIXED_DETS)
torch.cuda.synchronize()
im = pipe.infer_model
N = pipe.N
st = {}

def t_graph(name, fn, iters=10):
    g, o = pipe._capture(fn)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    st[name] = a.elapsed_time(b) / iters
    return o

with torch.no_grad():
    feat, table, ncls = t_graph("erfnet (to the 16-ch decoder map)", lambda: pipe.seg_model.forward_features_nhwc(pipe.rgbs.view(B * 3, 288, 256, 3)))
    t_graph("paint (fused seg head) + stack", lambda: (ops.paint_deconv_batched(pipe.lidar, feat, ncls, table, pipe._cams, 4, pipe.cur, (288, 256)),
                                                      ops.stack_jobs(pipe.jobs_dev, B * 3, N, 8, 3)))
    cv = t_graph("pillars (h16 canvas)", lambda: im.lidar_model.point_pillar_net.forward_nhwc(pipe.stacked, [3 * N] * B, canvas16=True))
    feats = t_graph("backbone", lambda: im.lidar_model.backbone.forward_nhwc(cv))
    heads = t_graph("heads", lambda: im.lidar_model.heads_nhwc(feats))
    t_graph("peaks", lambda: ops.det_peaks(heads[0], heads[1], heads[2]))
    wide = pipe.rgbs.permute(0, 2, 1, 3, 4).reshape(B, 288, 768, 3).permute(0, 3, 1, 2).float().contiguous(memory_format=torch.channels_last)
    tel = pipe.tels.permute(0, 3, 1, 2).float().contiguous(memory_format=torch.channels_last)
    t_graph("brake", lambda: pipe._brake())
    K = 3 * B
    g, o, s2 = pipe._g2[-(-K // pipe.K_BUCKET) * pipe.K_BUCKET]
    up = im.uniplanner
    fn = feats.permute(0, 3, 1, 2)
    crops = t_graph("crop", lambda: up.crop_feature(fn, s2["locs"], s2["oris"], pixels_per_meter=2.0, crop_size=96, frame_idx=s2["fidx"]))
    cr = crops.to(up.lidar_conv_emb[0].conv1.weight.dtype)
    embd = t_graph("embed_resnet18", lambda: up.lidar_conv_emb(cr)).float()
    cast = t_graph("cast_grus", lambda: up.cast(embd))
    t_graph("plan_grus", lambda: up.plan(embd[K:], pipe.nxps, cast_locs=cast[K:], pixels_per_meter=4, crop_size=192))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        pipe._g1.replay(); g.replay()
    b.record(); torch.cuda.synchronize()
    st["G1+G2 replay"] = a.elapsed_time(b) / 10
tot = sum(v for k, v in st.items() if k != "G1+G2 replay")
for k, v in st.items():
    print(f"{k:16s} {v:8.3f} ms  {v / B * 1e3:8.1f} us/frame  {100 * v / tot:5.1f}%")
print("sum of stages", tot, "B", B)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"B": B, "stages_ms": st}, open(f"gpurun_out/stage_times_b{B}.json", "w"), indent=1)
