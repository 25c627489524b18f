"""Step time of train_lidar (32 samples, fp32) under cheap host-side variants: fused Adam, channels_last, cuDNN autotune."""
import os, sys, time, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lav_b200.train import LAVTrainer, synthetic_train_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
(seg, lid0, uni0, bra), _ = bench.build_models()
batch = synthetic_train_batch(B, dev)

def run(tag, fused=False, cl=False, bench_=False, amp=False):
    torch.backends.cudnn.benchmark = bench_
    lid, uni = copy.deepcopy(lid0).to(dev), copy.deepcopy(uni0).to(dev)
    if cl:
        lid.to(memory_format=torch.channels_last); uni.to(memory_format=torch.channels_last)
    tr = LAVTrainer(lid, uni, device=dev, amp=amp)
    if fused:
        tr.optim = torch.optim.Adam(tr.params, lr=3e-4, fused=True)
    for _ in range(4):
        loss, _p = tr.train_lidar(*batch)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(5):
        loss, _p = tr.train_lidar(*batch)
    torch.cuda.synchronize()
    print(f"{tag:40s} {(time.time() - t0) / 5 * 1e3:7.1f} ms/step   loss {float(loss):.4f}", flush=True)
    tr.reducer.close()

import lav_b200.heads as Hd
import lav_b200.train as Tr
run("product (crop kernel + fused heads)")
Hd.TRAIN_CROP_KERNEL = False
run("  grid_sample crops")
Hd.TRAIN_CROP_KERNEL = True
Tr.FUSE_HEADS_TRAIN = False
run("  per-head first layers")
Tr.FUSE_HEADS_TRAIN = True
run("channels_last", cl=True)
run("fused adam + cudnn.benchmark", fused=True, bench_=True)
run("bf16 autocast (LiDAR model only)", amp=True)
