"""torch.profiler KERNEL table of one train_lidar step (32 samples): device kernels only (no aten:: op rows), their launch
count, summed duration, and the GPU-busy fraction of the step's wall time."""
import os, sys, time, collections
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
t0 = time.time()
for _ in range(3):
    tr.train_lidar(*batch)
torch.cuda.synchronize()
print(f"un-profiled step: {(time.time() - t0) / 3 * 1e3:.1f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.train_lidar(*batch)
    torch.cuda.synchronize()
ks = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ks:
    a = agg[e.name]; a[0] += 1; a[1] += e.device_time
tot = sum(v[1] for v in agg.values())
iv = sorted((e.time_range.start, e.time_range.end) for e in ks)
busy, cur_s, cur_e = 0.0, None, None
for s, e in iv:
    if cur_e is None or s > cur_e:
        if cur_e is not None: busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += (cur_e - cur_s) if cur_e is not None else 0
span = iv[-1][1] - iv[0][0]
print(f"{len(ks)} device kernels/memcpys, summed {tot / 1e3:.1f} ms, union busy {busy / 1e3:.1f} ms over a span of {span / 1e3:.1f} ms")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{us / 1e3:9.2f} ms  {100 * us / tot:5.1f}%  n={n:4d}  {name[:120]}")
# per autograd op class: where the time goes by aten op (CPU-side key, device time attributed)
ops_ = [e for e in prof.key_averages() if e.key.startswith("aten::") or "Backward" in e.key or e.key.startswith("_")]
print("--- by op (self device time)")
for e in sorted(ops_, key=lambda e: -e.self_device_time_total)[:30]:
    print(f"{e.self_device_time_total / 1e3:9.2f} ms  n={e.count:4d}  {e.key[:100]}")
