#!/bin/bash
mkdir -p gpurun_out
L="22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 22,16,56,56,45,64,3,1,1,1,1,1,1,0,0 22,8,28,28,288,128,3,1,1,1,1,1,1,0,0 22,4,14,14,576,256,3,1,1,1,1,1,1,0,0 22,16,56,56,64,64,1,3,3,1,1,1,0,1,1"
out=gpurun_out/r18_ab.txt; : > $out
echo "== default" >> $out; timeout 200 python tools/bench_conv.py $L >> $out 2>&1
echo "== ZSV_HALO_2CTA=1" >> $out; ZSV_HALO_2CTA=1 timeout 200 python tools/bench_conv.py $L >> $out 2>&1
cat $out
for m in def force def force; do
  if [ $m = def ]; then unset ZSV_HALO_2CTA; else export ZSV_HALO_2CTA=1; fi
  python bench.py --no-cpu-baseline > gpurun_out/r18_bench_$m.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/r18_bench_$m.json').read().strip().splitlines()[-1]); print('$m', d['value'], d['ms_per_step'])"
done
