#!/bin/bash
# A/B of library variants on ONE box: isolated layer-1 convolutions (graph-replayed, tools/bench_conv.py)
specs="22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 22,16,56,56,45,64,3,1,1,1,1,1,1,0,0 22,8,28,28,128,288,1,3,3,1,1,1,0,1,1 22,8,28,28,288,128,3,1,1,1,1,1,1,0,0"
for v in "$@"; do
  echo "== $v"
  ZSV_LIB_PATH=zeroshotvideoclassification_b200/build/variants/$v.so timeout 300 python tools/bench_conv.py $specs 2>&1 | grep -v "^\[zsv\]"
done
