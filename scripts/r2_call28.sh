#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call28.log
: > $LOG
echo "=== frame tests (cast kernel v2)" >> $LOG
timeout 900 python -m pytest tests/test_gpu_frame.py -m gpu -q -x 2>&1 | tail -3 >> $LOG
echo "=== stage times B=32" >> $LOG
timeout 600 python scripts/stage_times.py 32 2>&1 | tail -14 >> $LOG
tail -30 $LOG
