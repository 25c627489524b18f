#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call20.log
: > $LOG
echo "=== pair tests" >> $LOG
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "pair" 2>&1 | tail -3 >> $LOG
echo "=== trace" >> $LOG
timeout 300 python scripts/pair_trace.py 32 >> $LOG 2>&1
tail -30 $LOG
