#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/r20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r20_pytest.log
tail -5 gpurun_out/r20_pytest.log
L="22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 22,16,56,56,64,64,1,3,3,1,1,1,0,1,1"
out=gpurun_out/r20_ab.txt; : > $out
echo "== ZSV_HALO_WSHIFT=0" >> $out; ZSV_HALO_WSHIFT=0 timeout 200 python tools/bench_conv.py $L >> $out 2>&1
echo "== default (wshift)" >> $out; timeout 200 python tools/bench_conv.py $L >> $out 2>&1
cat $out
