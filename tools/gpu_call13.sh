#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nearest.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/c13_pytest.log
for G in 2 3; do for C in 101 200; do
ZSV_DEBUG_NEAREST_G=$G timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed -k regex:nearest -s 2 -c 1 python tools/ncu_nearest.py 10000 $C 2>&1 | grep -E "nearest_kernel|gpu__time|fp64" | tr '\n' ' '; echo " G=$G C=$C"
done; done 2>&1 | tee gpurun_out/c13_g_sweep.txt
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:nearest -s 2 -c 1 python tools/ncu_nearest.py 22 664 2>&1 | grep -E "gpu__time" | tee -a gpurun_out/c13_g_sweep.txt
