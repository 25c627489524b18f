#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call19.log
: > $LOG
timeout 300 python scripts/pair_trace.py 32 >> $LOG 2>&1
tail -30 $LOG
