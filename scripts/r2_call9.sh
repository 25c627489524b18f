#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call9.log
: > $LOG
echo "=== pytest" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 --tb=line 2>&1 | tail -15 >> $LOG
echo "=== pillar A/B" >> $LOG
timeout 300 python scripts/pillar_ab.py 32 2>&1 | grep "120k" >> $LOG
echo "=== smoke" >> $LOG
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 >> $LOG
echo "=== stage times" >> $LOG
timeout 300 python scripts/stage_times.py 32 2>&1 | tail -16 >> $LOG
tail -70 $LOG
