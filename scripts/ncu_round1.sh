set -x
NCU="ncu --clock-control none --profile-from-start off"
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r01_pillar_v3_launches.csv python scripts/pillar_layers.py 16 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:pillar_encode -c 1 -f -o gpurun_out/pillar_v3 python scripts/pillar_layers.py 16 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:stem7x7 -c 1 -f -o gpurun_out/stem_v2 python scripts/stem_layers.py 32 > /dev/null 2>&1
timeout 500 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r01_bench_b8_launches.csv -c 4000 python bench.py --steps 1 --warmup 1 --batch 8 --pipelines 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -8
