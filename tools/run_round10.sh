#!/bin/bash
mkdir -p gpurun_out
L="22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 22,8,28,28,128,288,1,3,3,1,1,1,0,1,1"
out=gpurun_out/r10_attr.txt; : > $out
for e in 0 1 2 3 4 7 8 15; do echo "== ZSV_DEBUG_EPI=$e" >> $out; ZSV_DEBUG_EPI=$e python tools/bench_conv.py $L >> $out 2>&1; done
echo "== STAGES=2" >> $out; ZSV_DEBUG_STAGES=2 python tools/bench_conv.py $L >> $out 2>&1
echo "== STAGES=3" >> $out; ZSV_DEBUG_STAGES=3 python tools/bench_conv.py $L >> $out 2>&1
echo "== NO_HALO" >> $out; ZSV_DEBUG_NO_HALO=1 python tools/bench_conv.py $L >> $out 2>&1
cat $out
