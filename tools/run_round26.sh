#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/r26_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r26_pytest.log
tail -3 gpurun_out/r26_pytest.log
L="22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 22,16,56,56,45,64,3,1,1,1,1,1,1,0,0 22,16,56,56,64,144,1,3,3,1,1,1,0,1,1"
out=gpurun_out/r26_ab.txt; : > $out
echo "== previous build" >> $out; ZSV_LIB_PATH=build/ab/libzsv_head.so timeout 200 python tools/bench_conv.py $L >> $out 2>&1
echo "== new" >> $out; timeout 200 python tools/bench_conv.py $L >> $out 2>&1
cat $out
