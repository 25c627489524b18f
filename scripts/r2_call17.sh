#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call17.log
: > $LOG
echo "=== pair kernel tests" >> $LOG
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "pair" 2>&1 | tail -5 >> $LOG
echo "=== conv + frame + config-size tests" >> $LOG
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_frame.py tests/test_gpu_config_sizes.py -m gpu -q 2>&1 | tail -8 >> $LOG
echo "=== stage times B=32" >> $LOG
timeout 600 python scripts/stage_times.py 32 2>&1 | tail -20 >> $LOG
echo "=== bench (no train)" >> $LOG
timeout 600 python bench.py --no-gpu-reference --no-cpu-baseline --no-train > gpurun_out/r2_bench_p.json 2> gpurun_out/r2_bench_p.err
python - >> $LOG <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_p.json').read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "parity", d.get("parity"), "roofline", d.get("roofline"))
except Exception as e:
    print("parse failed", e)
PY
echo "=== train profile (top)" >> $LOG
timeout 300 python scripts/train_profile.py 32 2>&1 | grep -v Warning | head -24 >> $LOG
tail -120 $LOG
