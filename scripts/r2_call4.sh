#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call4.log
: > $LOG
echo "=== pytest (failed ones)" >> $LOG
timeout 900 python -m pytest tests -q -m gpu --timeout 400 -k "config or lidar_model_f16 or infer_model or frame_pipeline or static_pipeline or paint_from_decoder or erf_nb16 or roof or brake_real or folded or train_losses" 2>&1 | tail -40 >> $LOG
echo "=== smoke" >> $LOG
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 >> $LOG
echo "=== ncu pillar tiled" >> $LOG
NCU="ncu --clock-control none --profile-from-start off"
timeout 200 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_pillar_tiled_launches.csv python scripts/pillar_layers.py 32 > /dev/null 2>&1
timeout 200 $NCU --set full --import-source on -k regex:pillar_tile_encode -c 1 -f -o gpurun_out/r2_pillar_tiled python scripts/pillar_layers.py 32 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_pillar_tiled_launches.csv | cut -d, -f5,9,13- | tail -12 >> $LOG
echo "=== erfnet time" >> $LOG
timeout 200 python scripts/r2_explore.py 32 2>&1 | head -1 >> $LOG
tail -80 $LOG
