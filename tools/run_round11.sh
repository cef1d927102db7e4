#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/r11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r11_pytest.log
tail -15 gpurun_out/r11_pytest.log
L="22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 22,8,28,28,128,288,1,3,3,1,1,1,0,1,1 22,16,56,56,45,64,3,1,1,1,1,1,1,0,0"
out=gpurun_out/r11_ab.txt; : > $out
for m in 0 auto 1; do echo "== ZSV_HALO_2CTA=$m" >> $out
  if [ $m = auto ]; then timeout 200 python tools/bench_conv.py $L >> $out 2>&1; else ZSV_HALO_2CTA=$m timeout 200 python tools/bench_conv.py $L >> $out 2>&1; fi; done
echo "== auto EPI=15" >> $out; ZSV_DEBUG_EPI=15 timeout 200 python tools/bench_conv.py $L >> $out 2>&1
echo "== auto EPI=7" >> $out; ZSV_DEBUG_EPI=7 timeout 200 python tools/bench_conv.py $L >> $out 2>&1
cat $out
