#!/bin/bash
# round-end validation on one B200: full GPU test suite, smoke, DRAM-traffic refresh, the bench record, launch list + ncu captures.
mkdir -p gpurun_out
LOG=gpurun_out/final_validation.log
: > $LOG
echo "=== pytest -m gpu" >> $LOG
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 >> $LOG
echo "=== smoke" >> $LOG
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $LOG
echo "=== traffic" >> $LOG
timeout 1500 python scripts/ncu_traffic.py 32 > gpurun_out/r02_traffic_run.log 2>&1
cp profiles/r02_traffic.json gpurun_out/r02_traffic.json 2>/dev/null
tail -c 600 gpurun_out/r02_traffic_run.log >> $LOG
echo "=== bench" >> $LOG
timeout 1200 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 300 gpurun_out/r02_bench.err >> $LOG
python - >> $LOG <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","roofline","roofline_pillar","parity","latency_b1","gpu_reference","train","cpu_baseline","clocks","gpu_launches"):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print("bench parse failed", e)
PY
echo "=== stage times" >> $LOG
timeout 600 python scripts/stage_times.py 32 2>&1 | tail -14 >> $LOG
cp gpurun_out/stage_times_b32.json gpurun_out/r02_stage_times_b32.json 2>/dev/null
echo "=== ncu launch list + captures" >> $LOG
timeout 500 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_bench_b16_launches.csv -c 6000 \
    python bench.py --steps 1 --warmup 1 --batch 16 --pipelines 1 --no-cpu-baseline --no-train --no-gpu-reference --no-graphs > gpurun_out/r02_bench_under_ncu.log 2>&1
NCU="ncu --clock-control none --profile-from-start off"
timeout 300 $NCU --set full --import-source on -k regex:conv_pair_umma -c 2 -f -o gpurun_out/r02_erf_pair python scripts/erfnet_range.py 32 > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02_erfnet_b32_launches.csv python scripts/erfnet_range.py 32 > /dev/null 2>&1
timeout 300 python scripts/pair_trace.py 32 > gpurun_out/r02_pair_trace.txt 2>&1
ls -la gpurun_out | tail -15 >> $LOG
tail -80 $LOG
