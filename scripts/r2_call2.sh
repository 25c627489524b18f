#!/bin/bash
# round 2, call 2: whole GPU suite on the f16 build, error table, sub-batch exploration, ncu of the ERFNet kernels
mkdir -p gpurun_out
LOG=gpurun_out/r2_call2.log
: > $LOG
echo "=== pytest -m gpu" >> $LOG
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -25 >> $LOG
echo "=== h16 error" >> $LOG
timeout 200 python scripts/h16_error.py 2>&1 | tail -12 >> $LOG
echo "=== explore" >> $LOG
timeout 300 python scripts/r2_explore.py 32 2>&1 | tail -12 >> $LOG
echo "=== ncu erfnet" >> $LOG
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --csv --log-file gpurun_out/r2_erfnet_b32_launches.csv python scripts/erfnet_profile.py 32 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:conv_pair_umma -s 12 -c 1 -f -o gpurun_out/r2_pair python scripts/erfnet_profile.py 32 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:conv_umma_kernel -s 20 -c 1 -f -o gpurun_out/r2_umma128 python scripts/erfnet_profile.py 32 > /dev/null 2>&1
ls -la gpurun_out | tail -6 >> $LOG
tail -70 $LOG
