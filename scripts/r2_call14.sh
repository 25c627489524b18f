#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call14.log
: > $LOG
echo "=== train profile" >> $LOG
timeout 300 python scripts/train_profile.py 32 2>&1 | tail -30 >> $LOG
echo "=== 2-GPU bench (torchrun)" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 600 gpurun_out/r2_bench_n2.err >> $LOG
python - >> $LOG <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
    for k in ("value","n_gpus","ms_per_step","e2e","train"):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e:
    print("bench parse failed", e)
PY
echo "=== reference arm under torchrun" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-600 >> $LOG
tail -60 $LOG
