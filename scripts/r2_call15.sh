#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call15.log
: > $LOG
echo "=== quick tests (static pipeline / copy stream, ddp reducer path)" >> $LOG
timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -5 >> $LOG
echo "=== bench copy stream off" >> $LOG
LAVB_COPY_STREAM=0 timeout 600 python bench.py --no-gpu-reference --no-cpu-baseline --no-train > gpurun_out/r2_bench_cs0.json 2> gpurun_out/r2_bench_cs0.err
echo "=== bench copy stream on (+train)" >> $LOG
timeout 600 python bench.py --no-gpu-reference --no-cpu-baseline > gpurun_out/r2_bench_cs1.json 2> gpurun_out/r2_bench_cs1.err
python - >> $LOG <<'PY'
import json
for f in ("cs0","cs1"):
    try:
        d=json.loads(open(f'gpurun_out/r2_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "train", (d.get("train") or {}).get("value"), (d.get("train") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "parse failed", e)
PY
tail -c 400 gpurun_out/r2_bench_cs1.err >> $LOG
echo "=== train profile" >> $LOG
timeout 300 python scripts/train_profile.py 32 2>&1 | grep -v Warning | tail -85 >> $LOG
echo "=== train variants" >> $LOG
timeout 600 python scripts/train_variants.py 32 2>&1 | grep "ms/step\|Error\|error" >> $LOG
tail -150 $LOG
