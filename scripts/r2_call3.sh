#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call3.log
: > $LOG
echo "=== pytest -m gpu" >> $LOG
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 2>&1 | tail -40 >> $LOG
echo "=== h16 error" >> $LOG
timeout 200 python scripts/h16_error.py 2>&1 | tail -12 >> $LOG
echo "=== pillar A/B" >> $LOG
timeout 300 python scripts/pillar_ab.py 32 2>&1 | tail -16 >> $LOG
echo "=== smoke" >> $LOG
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 >> $LOG
tail -90 $LOG
