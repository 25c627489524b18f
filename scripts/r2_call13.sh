#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call13.log
: > $LOG
echo "=== pytest" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 --tb=line 2>&1 | tail -12 >> $LOG
echo "=== erfnet" >> $LOG
timeout 200 python scripts/r2_explore.py 32 2>&1 | head -1 >> $LOG
timeout 200 python scripts/erfnet_profile.py 32 2>&1 | grep "us  n=" | head -8 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
tail -c 800 gpurun_out/r2_bench_c.err >> $LOG
python - >> $LOG <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_c.json').read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","roofline_pillar","parity","latency_b1","train"):
        print(k, json.dumps(d.get(k))[:700])
    print("gpu_reference", {k: v for k, v in (d.get("gpu_reference") or {}).items() if k not in ("what", "driven_like", "outputs")})
except Exception as e:
    print("bench parse failed", e)
PY
tail -50 $LOG
