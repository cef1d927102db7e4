#!/bin/bash
# sensitivity of one convolution to the kernels' tuning switches, one box: tools/ab_env.sh VARIANT SPEC "ENV=.. ENV=.." ...
v=$1; spec=$2; shift 2
for e in "$@"; do
  echo "== $v [$e]"
  env $e ZSV_LIB_PATH=zeroshotvideoclassification_b200/build/variants/$v.so timeout 300 python tools/bench_conv.py $spec 2>&1 | grep -v "wgrad pair plan" | cut -c49-200
done
