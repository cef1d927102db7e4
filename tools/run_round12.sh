#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r12_pytest.log
tail -4 gpurun_out/r12_pytest.log
for m in 0 auto 0 auto; do
  if [ $m = auto ]; then python bench.py --no-cpu-baseline > gpurun_out/r12_bench_$m.json 2>/dev/null; else ZSV_HALO_2CTA=0 python bench.py --no-cpu-baseline > gpurun_out/r12_bench_$m.json 2>/dev/null; fi
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r12_bench_$m.json').read().strip().splitlines()[-1]); print('$m', d['value'], d['ms_per_step'])"
done
L="22,16,56,56,64,144,1,3,3,1,1,1,0,1,1"
python tools/ncu_one.py $L dgrad_fused && timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_halo -s 2 -c 1 -o gpurun_out/r12_halo_pair_dgrad_fused -f python tools/ncu_one.py $L dgrad_fused > gpurun_out/r12_ncu.log 2>&1
tail -2 gpurun_out/r12_ncu.log
