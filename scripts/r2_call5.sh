#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call5.log
: > $LOG
timeout 900 python -m pytest tests -q -m gpu --timeout 400 --tb=line -k "config or lidar_model_f16 or infer_model or frame_pipeline or static_pipeline or paint_from_decoder" 2>&1 | tail -30 >> $LOG
echo "=== smoke" >> $LOG
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 >> $LOG
echo "=== erfnet kernels" >> $LOG
timeout 200 python scripts/erfnet_profile.py 32 2>&1 | tail -14 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 3000 gpurun_out/r2_bench_a.err >> $LOG
python - >> $LOG <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_a.json').read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","roofline","roofline_pillar","cpu_baseline","parity","latency_b1","gpu_reference","train","clocks"):
        print(k, json.dumps(d.get(k))[:600])
except Exception as e:
    print("bench parse failed", e)
PY
tail -70 $LOG
