#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r21_pytest.log
tail -5 gpurun_out/r21_pytest.log
L="22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 22,16,56,56,64,64,1,3,3,1,1,1,0,1,1"
out=gpurun_out/r21_ab.txt; : > $out
echo "== ZSV_HALO_WSHIFT=0" >> $out; ZSV_HALO_WSHIFT=0 timeout 200 python tools/bench_conv.py $L >> $out 2>&1
echo "== default (wshift)" >> $out; ZSV_DEBUG_PLAN=1 timeout 200 python tools/bench_conv.py $L >> $out 2>&1
grep -v "halo plan" $out; grep "halo plan" $out | sort | uniq -c
for m in off on off on; do
  if [ $m = on ]; then unset ZSV_HALO_WSHIFT; else export ZSV_HALO_WSHIFT=0; fi
  python bench.py --no-cpu-baseline > gpurun_out/r21_bench_$m.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/r21_bench_$m.json').read().strip().splitlines()[-1]); print('$m', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
