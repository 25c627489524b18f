"""Round-2 entry point for the prepared (off-by-default) kernels: flag off vs flag on, numerics + CUDA-graph replay time.
  python scripts/experimental_check.py [B] [pairs] [halo] [epi16] [gru]        (no selector = all four)
    pairs : erfnet.FUSE_PAIRS   fused (3x1 -> 1x3) tcgen05 pairs            -> ERFNet
    halo  : layers.USE_HALO     halo-patch tcgen05 convolution               -> ERFNet, BEV backbone
    epi16 : layers.USE_EPI16    narrow layers, 2 CTAs/SM x 8 epilogue warps  -> ERFNet, BEV backbone
    gru   : heads.GRU_KERNEL    cluster-persistent plan GRU                  -> planner roll-out
    trunk : ResNet-18 trunk of the brake model on the tcgen05 kernels (resnet_umma.py), with and without USE_HALO / USE_EPI16
            (in round 1 that trunk was 5-20 % slower than BN-folded cuDNN; the narrow-layer switches may flip it)
Run each selector in its own process under `timeout` (scripts/round2_first_call.sh): a hang then costs one item."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lav_b200 import erfnet, heads, layers, ops, synth
from tests import util

args = sys.argv[1:]
B = int(args[0]) if args and args[0].isdigit() else 32
WHICH = {a for a in args if not a.isdigit()} or {"pairs", "halo", "epi16", "gru"}
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


def compare(label, module, flag, fn):
    res = {}
    for on in (False, True):
        setattr(module, flag, on)
        ms, out = graph_time(fn)
        res[on] = (ms, out.float().clone())
    setattr(module, flag, False)
    d = (res[True][1] - res[False][1]).abs().max().item() / max(res[False][1].abs().max().item(), 1e-30)
    print(f"{label}: {flag} off {res[False][0]:.3f} ms, on {res[True][0]:.3f} ms ({res[False][0] / res[True][0]:.2f}x), max-norm diff {d:.2e}", flush=True)


with torch.no_grad():
    if WHICH & {"pairs", "halo", "epi16"}:
        seg, _ = util.seg_model(dev)
        seg.set_precision("f16")
        rgb = synth.rgb_frames().to(dev).repeat(B, 1, 1, 1)
        lid, _ = util.lidar_model(dev)
        lid.set_precision("f16")
        canvas = (torch.randn(B, 320, 320, 128, device=dev) * 0.3).to(ops.h16())       # [hi | lo] split canvas
        run_seg, run_bb = (lambda: seg.forward_nhwc(rgb)), (lambda: lid.backbone.forward_nhwc(canvas))
        if "epi16" in WHICH:
            compare(f"ERFNet {3 * B} images", layers, "USE_EPI16", run_seg)
            compare(f"BEV backbone {B} frames", layers, "USE_EPI16", run_bb)
        if "halo" in WHICH:
            compare(f"ERFNet {3 * B} images", layers, "USE_HALO", run_seg)
            compare(f"BEV backbone {B} frames", layers, "USE_HALO", run_bb)
        if "pairs" in WHICH:
            compare(f"ERFNet {3 * B} images", erfnet, "FUSE_PAIRS", run_seg)
    if "trunk" in WHICH:
        import bench
        (_, _, _, bra), _ = bench.build_models()
        bra = bra.to(dev).eval()
        bra.conv_backbone.to(ops.h16()).to(memory_format=torch.channels_last)
        bra.attn1.to(ops.h16()); bra.attn2.to(ops.h16())
        g = torch.Generator().manual_seed(1)
        rgbs = torch.randint(0, 256, (B, 3, 288, 256, 3), generator=g, dtype=torch.uint8).to(dev)
        tel = torch.randint(0, 256, (B, 192, 480, 3), generator=g, dtype=torch.uint8).to(dev)
        base = None
        for umma_trunk, halo, epi16 in ((False, False, False), (True, False, False), (True, True, False), (True, False, True)):
            bra.conv_backbone.use_umma_trunk = umma_trunk
            layers.USE_HALO, layers.USE_EPI16 = halo, epi16
            ms, out = graph_time(lambda: bra.forward_u8(rgbs, tel))
            base = out.float().clone() if base is None else base
            d = (out.float() - base).abs().max().item()
            print(f"brake model {B} frames: tcgen05 trunk {umma_trunk}, halo {halo}, epi16 {epi16}: {ms:.3f} ms, max |diff| vs cuDNN trunk {d:.2e}", flush=True)
        bra.conv_backbone.use_umma_trunk = False
        layers.USE_HALO = layers.USE_EPI16 = False
    if "gru" in WHICH:
        gru = torch.nn.GRU(4, 512, batch_first=True).to(dev)
        mlp = torch.nn.Linear(512, 2).to(dev)
        embd = torch.randn(B, 512, device=dev) * 0.5
        nxp = torch.tensor([[0.0, -20.0]] * B, device=dev)
        cast = torch.randn(B, 6, 20, 2, device=dev) * 0.1
        compare(f"plan roll-out {6 * B} sequences x 20 steps x 5 iterations", heads, "GRU_KERNEL",
                lambda: heads._plan_rollout(gru, mlp, 6, 20, 5, embd, nxp, cast, 4, 192))
