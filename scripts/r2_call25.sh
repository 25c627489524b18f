#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call25.log
: > $LOG
echo "=== gru + conv tests" >> $LOG
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q 2>&1 | tail -3 >> $LOG
echo "=== frame + config tests" >> $LOG
timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_gpu_config_sizes.py -m gpu -q 2>&1 | tail -3 >> $LOG
echo "=== stage times B=32 (cuDNN trunks)" >> $LOG
timeout 600 python scripts/stage_times.py 32 2>&1 | tail -14 >> $LOG
echo "=== stage times B=32 (UMMA trunks)" >> $LOG
LAVB_UMMA_TRUNKS=1 timeout 600 python scripts/stage_times.py 32 2>&1 | tail -14 >> $LOG
tail -60 $LOG
