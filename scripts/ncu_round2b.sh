#!/bin/bash
# second set of round-2 ncu evidence: the fused heads conv and a 64-channel backbone layer after the coalesced-store epilogue,
# the crop backward kernel of the training step, the cast kernel.  Summaries: scripts/summarize_ncu.py rep <file> -> profiles/.
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 300 $NCU --set full --import-source on -k regex:conv_umma_kernel -c 1 -f -o gpurun_out/r02_heads_conv python scripts/heads_conv_layer.py 32 > /dev/null 2>&1
timeout 300 ncu --clock-control none --set full --import-source on -k regex:conv_umma_kernel -s 16 -c 1 -f -o gpurun_out/r02_bb64_conv python scripts/umma_layers.py 32 > gpurun_out/r02_umma_layers.txt 2>&1
timeout 300 $NCU --set full --import-source on -k regex:crop_bwd -c 1 -f -o gpurun_out/r02_crop_bwd python scripts/crop_bwd_layer.py > gpurun_out/r02_crop_bwd.txt 2>&1
timeout 300 python scripts/crop_bwd_layer.py > gpurun_out/r02_crop_bwd.txt 2>&1
timeout 300 python scripts/umma_layers.py 32 > gpurun_out/r02_umma_layers.txt 2>&1
ls -la gpurun_out/r02_heads_conv.ncu-rep gpurun_out/r02_bb64_conv.ncu-rep gpurun_out/r02_crop_bwd.ncu-rep
cat gpurun_out/r02_crop_bwd.txt gpurun_out/r02_umma_layers.txt
