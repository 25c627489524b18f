#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call16.log
: > $LOG
echo "=== tests: crop bwd, train, dropin" >> $LOG
timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_gpu_train.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -8 >> $LOG
echo "=== train variants" >> $LOG
timeout 600 python scripts/train_variants.py 32 2>&1 | grep "ms/step\|Error\|error" >> $LOG
echo "=== train profile" >> $LOG
timeout 300 python scripts/train_profile.py 32 2>&1 | grep -v Warning | tail -70 >> $LOG
tail -150 $LOG
