#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call29.log
: > $LOG
echo "=== 2-GPU bench (torchrun)" >> $LOG
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 400 gpurun_out/r02_bench_n2.err >> $LOG
python - >> $LOG <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
    for k in ("value","n_gpus","ms_per_step","e2e","train"):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e:
    print("bench parse failed", e)
PY
tail -20 $LOG
