#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call18.log
: > $LOG
echo "=== pipelines / batch sweep" >> $LOG
for cfg in "64 2" "64 4" "96 3" "128 4" "128 2"; do
  set -- $cfg
  timeout 400 python bench.py --batch $1 --pipelines $2 --steps 20 --no-gpu-reference --no-cpu-baseline --no-train > gpurun_out/r2_sweep.json 2> gpurun_out/r2_sweep.err
  python - "$1" "$2" >> $LOG <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2_sweep.json').read().strip().splitlines()[-1])
    print("B", sys.argv[1], "P", sys.argv[2], "value %.0f e2e %.0f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print("B", sys.argv[1], "P", sys.argv[2], "failed", e, open('gpurun_out/r2_sweep.err').read()[-300:])
PY
done
echo "=== ncu pair kernel (new epilogue)" >> $LOG
NCU="ncu --clock-control none --profile-from-start off"
timeout 300 $NCU --set full --import-source on -k regex:conv_pair_umma -c 4 -f -o gpurun_out/r02_erf_pair_v2 python scripts/erfnet_range.py 32 > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02_erfnet_b32_launches_v2.csv python scripts/erfnet_range.py 32 > /dev/null 2>&1
ls -la gpurun_out | grep "v2" >> $LOG
echo "=== train bench" >> $LOG
timeout 300 python scripts/train_variants.py 32 2>&1 | grep "ms/step\|Error\|error" | head -1 >> $LOG
tail -40 $LOG
