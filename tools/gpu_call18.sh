#!/bin/bash
# launch list of the current build (pair wgrad on / off) with the planner's decisions
mkdir -p gpurun_out
ZSV_DEBUG_PLAN=1 timeout 300 python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c18_quick.json 2> gpurun_out/c18_plan.err
sort -u gpurun_out/c18_plan.err | grep "wgrad pair plan" > gpurun_out/c18_wgrad_pair_plans.txt; wc -l gpurun_out/c18_wgrad_pair_plans.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c18_launches_pair.csv python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c18_ncu.log 2>&1; echo "ncu rc=$?"
ZSV_WGRAD_PAIR=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c18_launches_nopair.csv python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c18_ncu2.log 2>&1; echo "ncu rc=$?"
