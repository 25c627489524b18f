#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call10.log
: > $LOG
timeout 900 python -m pytest tests -q -m gpu --timeout 400 --tb=line -k "config3 or erf_stem or erfnet or static_pipeline or paint_from" 2>&1 | tail -12 >> $LOG
echo "=== stage times" >> $LOG
timeout 300 python scripts/stage_times.py 32 2>&1 | head -3 >> $LOG
tail -30 $LOG
