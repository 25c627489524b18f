"""torch.profiler kernel table of one train_lidar step (32 samples)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from lav_b200.train import LAVTrainer, synthetic_train_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
(seg, lid, uni, bra), _ = bench.build_models()
tr = LAVTrainer(lid.to(dev), uni.to(dev), device=dev, amp="amp" in sys.argv)
batch = synthetic_train_batch(B, dev)
for _ in range(3):
    tr.train_lidar(*batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.train_lidar(*batch)
    torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.self_device_time_total > 0]
tot = sum(e.self_device_time_total for e in evs)
print(f"total device time {tot / 1e3:.1f} ms over {sum(e.count for e in evs)} kernels")
for e in sorted(evs, key=lambda e: -e.self_device_time_total)[:28]:
    print(f"{e.self_device_time_total / 1e3:9.2f} ms  {100 * e.self_device_time_total / tot:5.1f}%  n={e.count:4d}  {e.key[:110]}")
