#!/bin/bash
mkdir -p gpurun_out
# fused BN-backward dgrad of conv 64->144 1x3x3 (dy has 144 channels, dx 64) and temporal fprop 144->64
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_halo -s 2 -c 1 -f -o gpurun_out/c15_dgrad_fused_64_144 python tools/ncu_one.py 22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 dgrad_fused > gpurun_out/c15_a.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_halo -s 2 -c 1 -f -o gpurun_out/c15_fprop_144_64 python tools/ncu_one.py 22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 fprop > gpurun_out/c15_b.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm -s 2 -c 1 -f -o gpurun_out/c15_dgrad_64_230 python tools/ncu_one.py 22,16,56,56,64,230,1,3,3,1,2,2,0,1,1 dgrad > gpurun_out/c15_c.log 2>&1; echo "rc=$?"
