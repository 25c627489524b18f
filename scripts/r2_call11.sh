#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call11.log
: > $LOG
timeout 900 python -m pytest tests -q -m gpu --timeout 400 --tb=line -k "data_pipeline or erf_stem or erfnet or conv_pair or frame_pipeline" 2>&1 | tail -8 >> $LOG
echo "=== erfnet pair variants" >> $LOG
LAVB_PAIR_TWO=0 timeout 200 python scripts/r2_explore.py 32 2>&1 | head -1 >> $LOG
LAVB_PAIR_TWO=1 timeout 200 python scripts/r2_explore.py 32 2>&1 | head -1 >> $LOG
timeout 200 python scripts/erfnet_profile.py 32 2>&1 | grep "us  n=" | head -9 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -c 1500 gpurun_out/r2_bench_b.err >> $LOG
python - >> $LOG <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","roofline","roofline_pillar","roofline_heads_conv","parity","latency_b1","train"):
        print(k, json.dumps(d.get(k))[:500])
    print("gpu_reference", {k: v for k, v in (d.get("gpu_reference") or {}).items() if k not in ("what", "driven_like", "outputs")})
except Exception as e:
    print("bench parse failed", e)
PY
tail -60 $LOG
