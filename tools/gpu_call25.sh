#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c25_pytest.log 2>&1; tail -2 gpurun_out/c25_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-extras --layer-table > gpurun_out/c25_bench.json 2> gpurun_out/c25_layer_table.txt; cut -c1-300 gpurun_out/c25_bench.json
ZSV_NYBUF=1 timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/c25_bench_ny1.json 2> /dev/null; cut -c1-300 gpurun_out/c25_bench_ny1.json
cap() { # name spec pass skip count
timeout 300 python tools/ncu_one.py $2 $3 > /dev/null 2>&1 || { echo "$1 plain run failed"; return; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm -s $4 -c $5 -f -o gpurun_out/c25_$1 python tools/ncu_one.py $2 $3 > gpurun_out/c25_$1.log 2>&1; echo "$1 rc=$?"; }
cap fprop_144_64 22,16,56,56,144,64,3,1,1,1,1,1,1,0,0 fprop 2 1
cap dgrad_fused_64_144 22,16,56,56,64,144,1,3,3,1,1,1,0,1,1 dgrad_fused 2 1
