#!/bin/bash
# round-2 ncu evidence: launch list of one bench tick (eager), --set full captures of the kernels DESIGN.md quotes.
# Usage under gpurun: bash scripts/ncu_round2.sh ; the .ncu-rep files and CSVs land in gpurun_out/, summaries are made locally
# with scripts/summarize_ncu.py and committed under profiles/.
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
# (1) launch list of the product tick: bench.py eager pass is the same G1 body the graphs capture
timeout 500 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_bench_b16_launches.csv -c 6000 \
    python bench.py --steps 1 --warmup 1 --batch 16 --pipelines 1 --no-cpu-baseline --no-train --no-gpu-reference --no-graphs > gpurun_out/r02_bench_under_ncu.log 2>&1
# (2) --set full captures
timeout 300 $NCU --set full --import-source on -k regex:pillar_encode_sorted -c 1 -f -o gpurun_out/r02_pillar_sorted python scripts/pillar_layers.py 32 > /dev/null 2>&1
LAVB_PILLAR_ENCODER=tiled timeout 300 $NCU --set full --import-source on -k regex:pillar_tile_encode_tc -c 1 -f -o gpurun_out/r02_pillar_tiled_tc python scripts/pillar_layers.py 32 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:conv_umma_kernel -c 1 -f -o gpurun_out/r02_heads_conv python scripts/heads_conv_layer.py 32 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:conv_pair_umma -c 2 -f -o gpurun_out/r02_erf_pair python scripts/erfnet_range.py 32 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:erf_nb16 -c 1 -f -o gpurun_out/r02_erf_nb16 python scripts/erfnet_range.py 32 > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02_erfnet_b32_launches.csv python scripts/erfnet_range.py 32 > /dev/null 2>&1
ls -la gpurun_out | tail -12
