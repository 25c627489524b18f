#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call7.log
: > $LOG
echo "=== pillar A/B" >> $LOG
timeout 300 python scripts/pillar_ab.py 32 2>&1 | tail -16 >> $LOG
echo "=== gru / plan debug" >> $LOG
timeout 200 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=line -k "gru_cluster" 2>&1 | tail -6 >> $LOG
timeout 200 python scripts/debug_plan.py 2>&1 | tail -5 >> $LOG
echo "=== new tests" >> $LOG
timeout 600 python -m pytest tests -q -m gpu --timeout 300 --tb=line -k "data_paint or config3 or pillar" 2>&1 | tail -10 >> $LOG
echo "=== traffic" >> $LOG
timeout 600 python scripts/ncu_traffic.py 32 2>&1 | tail -40 >> $LOG
cp profiles/r02_traffic.json gpurun_out/ 2>/dev/null
tail -100 $LOG
