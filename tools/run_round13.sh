#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/r13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r13_pytest.log
tail -4 gpurun_out/r13_pytest.log
L="22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 22,8,28,28,128,288,1,3,3,1,1,1,0,1,1 22,4,14,14,256,576,1,3,3,1,1,1,0,1,1"
out=gpurun_out/r13_ab.txt; : > $out
echo "== default" >> $out; timeout 200 python tools/bench_conv.py $L >> $out 2>&1
echo "== ZSV_2CTA=1" >> $out; ZSV_2CTA=1 timeout 200 python tools/bench_conv.py $L >> $out 2>&1
cat $out
for m in auto k2 auto k2; do
  if [ $m = auto ]; then python bench.py --no-cpu-baseline > gpurun_out/r13_bench_$m.json 2>/dev/null; else ZSV_2CTA=1 python bench.py --no-cpu-baseline > gpurun_out/r13_bench_$m.json 2>/dev/null; fi
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r13_bench_$m.json').read().strip().splitlines()[-1]); print('$m', d['value'], d['ms_per_step'])"
done
