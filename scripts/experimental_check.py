"""Round-2 entry point for the two prepared kernels: numerics + timing, flag off vs on, on one GPU.
  python scripts/experimental_check.py [B]
ERFNet with erfnet.FUSE_PAIRS (fused 3x1->1x3 tcgen05 pairs), ERFNet / BEV backbone with layers.USE_HALO (halo-patch conv)
and the planner roll-out with heads.GRU_KERNEL (cluster-persistent GRU).  Prints max-norm difference of the outputs and CUDA-graph replay times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lav_b200 import erfnet, heads, synth
from tests import util

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")


def graph_time(fn, iters=10):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out


with torch.no_grad():
    seg, _ = util.seg_model(dev)
    seg.set_precision("bf16")
    rgb = synth.rgb_frames().to(dev).repeat(B, 1, 1, 1)
    res = {}
    for flag in (False, True):
        erfnet.FUSE_PAIRS = flag
        ms, out = graph_time(lambda: seg.forward_nhwc(rgb))
        res[flag] = (ms, out.float().clone())
    d = (res[True][1] - res[False][1]).abs().max().item() / res[False][1].abs().max().item()
    print(f"ERFNet {3 * B} images: unfused {res[False][0]:.3f} ms, fused pairs {res[True][0]:.3f} ms, max-norm diff {d:.2e}")
    erfnet.FUSE_PAIRS = False

    # halo-patch conv kernel: ERFNet (its 3x1 layers) and the BEV backbone (3x3 layers)
    from lav_b200 import layers
    lid, _ = util.lidar_model(dev)
    lid.set_precision("bf16")
    canvas = (torch.randn(B, 320, 320, 128, device=dev) * 0.3).to(torch.bfloat16)      # [hi | lo] split canvas
    for name, fn in (("ERFNet", lambda: seg.forward_nhwc(rgb)), ("backbone", lambda: lid.backbone.forward_nhwc(canvas))):
        res = {}
        for flag in (False, True):
            layers.USE_HALO = flag
            ms, out = graph_time(fn)
            res[flag] = (ms, out.float().clone())
        d = (res[True][1] - res[False][1]).abs().max().item() / res[False][1].abs().max().item()
        print(f"{name}: per-tap tiles {res[False][0]:.3f} ms, halo patches {res[True][0]:.3f} ms, max-norm diff {d:.2e}")
    layers.USE_HALO = False
    for name, fn in (("ERFNet", lambda: seg.forward_nhwc(rgb)), ("backbone", lambda: lid.backbone.forward_nhwc(canvas))):
        res = {}
        for flag in (False, True):
            layers.USE_EPI16 = flag
            ms, out = graph_time(fn)
            res[flag] = (ms, out.float().clone())
        d = (res[True][1] - res[False][1]).abs().max().item()
        print(f"{name}: 2 CTAs x 4 epilogue warps {res[False][0]:.3f} ms, 2 CTAs x 8 epilogue warps {res[True][0]:.3f} ms, max |diff| {d:.2e}")
    layers.USE_EPI16 = False

    up, _ = util.uniplanner(dev) if hasattr(util, "uniplanner") else (None, None)
    gru = up.plan_gru if up is not None else torch.nn.GRU(4, 512, batch_first=True).to(dev)
    mlp = up.plan_mlp if up is not None else torch.nn.Linear(512, 2).to(dev)
    embd = torch.randn(B, 512, device=dev) * 0.5
    nxp = torch.tensor([[0.0, -20.0]] * B, device=dev)
    cast = torch.randn(B, 6, 20, 2, device=dev) * 0.1
    res = {}
    for flag in (False, True):
        heads.GRU_KERNEL = flag
        ms, out = graph_time(lambda: heads._plan_rollout(gru, mlp, 6, 20, 5, embd, nxp, cast, 4, 192))
        res[flag] = (ms, out.float().clone())
    d = (res[True][1] - res[False][1]).abs().max().item() / res[False][1].abs().max().item()
    print(f"plan roll-out {6 * B} sequences x 20 steps x 5 iterations: cuDNN {res[False][0]:.3f} ms, cluster kernel {res[True][0]:.3f} ms, "
          f"max-norm diff {d:.2e}")
    heads.GRU_KERNEL = False
