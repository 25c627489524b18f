#!/bin/bash
# First GPU call of round 2: validate and time the prepared (off-by-default) kernels, each isolated so that a hang in one
# (killed by `timeout`) does not cost the others.  Usage under gpurun:  bash scripts/round2_first_call.sh
export LAVB_EXPERIMENTAL=1
mkdir -p gpurun_out
LOG=gpurun_out/r2_first.log
: > $LOG
python -c "import torch; print(torch.cuda.get_device_name(0))" >> $LOG 2>&1
for k in umma16 conv_pair halo gru_cluster; do
  echo "=== test $k" >> $LOG
  timeout 120 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "$k" --timeout 50 --timeout-method=thread 2>&1 | tail -12 >> $LOG
done
for k in epi16 halo pairs gru trunk; do
  echo "=== timing $k" >> $LOG
  timeout 150 python scripts/experimental_check.py 32 $k 2>&1 | tail -6 >> $LOG
done
tail -80 $LOG
