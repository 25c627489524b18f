#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call6.log
: > $LOG
echo "=== pillar tests (tcgen05 tile encoder)" >> $LOG
timeout 600 python -m pytest tests/test_gpu_paint_pillar.py -q -m gpu --timeout 120 --tb=line -k "sorted_kernel or config2 or full_size" 2>&1 | tail -15 >> $LOG
timeout 300 python -m pytest tests/test_gpu_config_sizes.py -q -m gpu --timeout 200 --tb=line -k "config2" 2>&1 | tail -8 >> $LOG
echo "=== pillar A/B" >> $LOG
timeout 300 python scripts/pillar_ab.py 32 2>&1 | tail -16 >> $LOG
echo "=== pytest rest" >> $LOG
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 --tb=line 2>&1 | tail -25 >> $LOG
tail -70 $LOG
