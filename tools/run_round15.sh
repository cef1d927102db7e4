#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/r15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r15_pytest.log
tail -4 gpurun_out/r15_pytest.log
L="22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 22,16,56,56,45,64,3,1,1,1,1,1,1,0,0 22,8,28,28,288,128,3,1,1,1,1,1,1,0,0 22,8,28,28,128,288,1,3,3,1,1,1,0,1,1 22,8,28,28,128,230,1,3,3,1,1,1,0,1,1"
out=gpurun_out/r15_ab.txt; : > $out
echo "== HEAD lib" >> $out; ZSV_LIB_PATH=build/ab/libzsv_head.so timeout 200 python tools/bench_conv.py $L >> $out 2>&1
echo "== new lib" >> $out; ZSV_DEBUG_PLAN=1 timeout 200 python tools/bench_conv.py $L >> $out 2>&1
grep -v "halo plan" $out; grep "halo plan" $out | sort | uniq -c
for m in head new k2 head new k2; do
  case $m in head) export ZSV_LIB_PATH=build/ab/libzsv_head.so; unset ZSV_2CTA;; new) unset ZSV_LIB_PATH; unset ZSV_2CTA;; k2) unset ZSV_LIB_PATH; export ZSV_2CTA=1;; esac
  python bench.py --no-cpu-baseline > gpurun_out/r15_bench_$m.json 2>/dev/null
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r15_bench_$m.json').read().strip().splitlines()[-1]); print('$m', d['value'], d['ms_per_step'])"
done
