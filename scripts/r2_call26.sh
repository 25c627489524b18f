#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call26.log
: > $LOG
for sms in 148 140 132 124; do
  LAVB_NUM_SMS=$sms timeout 400 python bench.py --steps 20 --no-gpu-reference --no-cpu-baseline --no-train > gpurun_out/r2_sweep.json 2> gpurun_out/r2_sweep.err
  python - "$sms" >> $LOG <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2_sweep.json').read().strip().splitlines()[-1])
    print("SMs", sys.argv[1], "value %.0f e2e %.0f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print("SMs", sys.argv[1], "failed", e, open('gpurun_out/r2_sweep.err').read()[-300:])
PY
done
tail -10 $LOG
