#!/bin/bash
mkdir -p gpurun_out
L="22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 22,16,56,56,45,64,3,1,1,1,1,1,1,0,0 22,8,28,28,288,128,3,1,1,1,1,1,1,0,0"
out=gpurun_out/r14_ab.txt; : > $out
for i in 1 2; do
echo "== HEAD lib" >> $out; ZSV_LIB_PATH=build/ab/libzsv_head.so timeout 200 python tools/bench_conv.py $L >> $out 2>&1
echo "== new lib" >> $out; timeout 200 python tools/bench_conv.py $L >> $out 2>&1
done
cat $out
