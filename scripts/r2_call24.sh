#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call24.log
: > $LOG
echo "=== gru tests" >> $LOG
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "gru" 2>&1 | tail -5 >> $LOG
echo "=== frame + config tests" >> $LOG
timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_gpu_config_sizes.py tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -8 >> $LOG
echo "=== stage times B=32" >> $LOG
timeout 600 python scripts/stage_times.py 32 2>&1 | tail -14 >> $LOG
echo "=== smoke" >> $LOG
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> $LOG
tail -50 $LOG
