"""train_lidar step throughput (BASELINE config 4 shape: 32 samples per rank, NCCL gradient all-reduce).
    python scripts/train_bench.py [--batch 32] [--steps 5]        (or under torchrun for N ranks)"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from lav_b200.train import LAVTrainer, synthetic_train_batch

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=32); ap.add_argument("--steps", type=int, default=5); ap.add_argument("--amp", action="store_true", help="f16 autocast forwards (opt-in)")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
(seg, lid, uni, bra), _ = bench.build_models()
tr = LAVTrainer(lid.to(dev), uni.to(dev), device=dev, amp=args.amp)
batch = synthetic_train_batch(args.batch, dev, seed=2021 + rank)
for _ in range(2):
    tr.train_lidar(*batch)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss, parts = tr.train_lidar(*batch)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    n_par = sum(p.numel() for p in tr.params if p.requires_grad)
    print(json.dumps({"metric": "train_lidar_samples_per_s", "value": world * args.batch * args.steps / (float(ms) * 1e-3), "n_gpus": world,
                      "ms_per_step": float(ms) / args.steps, "per_rank_batch": args.batch, "allreduce_params": n_par,
                      "loss": float(loss), "max_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
if world > 1:
    dist.destroy_process_group()
