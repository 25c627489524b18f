#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call23.log
: > $LOG
echo "=== conv + frame + config tests" >> $LOG
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_frame.py tests/test_gpu_config_sizes.py -m gpu -q -x 2>&1 | tail -6 >> $LOG
echo "=== stage times B=32" >> $LOG
timeout 600 python scripts/stage_times.py 32 2>&1 | tail -14 >> $LOG
echo "=== bench" >> $LOG
timeout 600 python bench.py --no-gpu-reference --no-cpu-baseline --no-train > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err
python - >> $LOG <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_q.json').read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d.get("roofline"))
except Exception as e:
    print("parse failed", e)
PY
tail -40 $LOG
