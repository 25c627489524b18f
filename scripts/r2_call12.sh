#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call12.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_data_pipeline.py -q -m gpu --tb=line 2>&1 | tail -5 >> $LOG
bash scripts/ncu_round2.sh >> $LOG 2>&1
timeout 400 python scripts/ncu_traffic.py 32 > gpurun_out/r02_traffic_run.log 2>&1
cp profiles/r02_traffic.json gpurun_out/r02_traffic.json 2>/dev/null
tail -30 $LOG
