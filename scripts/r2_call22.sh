#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call22.log
: > $LOG
echo "=== stage times B=32" >> $LOG
timeout 600 python scripts/stage_times.py 32 2>&1 | tail -14 >> $LOG
echo "=== erfnet launches" >> $LOG
timeout 300 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_erfnet_b32_launches_v3.csv python scripts/erfnet_range.py 32 > /dev/null 2>&1
python - >> $LOG <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_erfnet_b32_launches_v3.csv')) if len(r)>14 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[14])/1e3
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{v[1]:9.1f} us n={v[0]:3d} {k}")
PY
echo "=== frame tests (erfnet parity)" >> $LOG
timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_config_sizes.py -m gpu -q 2>&1 | tail -3 >> $LOG
tail -40 $LOG
