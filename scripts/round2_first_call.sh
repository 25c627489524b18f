#!/bin/bash
# First GPU call of round 2: validate and time the prepared (off-by-default) kernels, each isolated so that a hang in one
# (killed by `timeout`) does not cost the others.  Usage under gpurun:  bash scripts/round2_first_call.sh 2>&1 | tail -60
export LAVB_EXPERIMENTAL=1
for k in umma16 conv_pair halo gru_cluster; do
  echo "=== test $k"
  timeout 150 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "$k" --timeout 60 --timeout-method=thread 2>&1 | tail -6
done
for k in epi16 halo pairs gru trunk; do
  echo "=== timing $k"
  timeout 200 python scripts/experimental_check.py 32 $k 2>&1 | tail -4
done
