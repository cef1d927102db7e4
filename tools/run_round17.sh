#!/bin/bash
mkdir -p gpurun_out
python bench.py --quick --no-graph --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r17_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r17_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r17_launches.csv > gpurun_out/r17_launch_summary.txt; head -30 gpurun_out/r17_launch_summary.txt
for i in 1 2; do python bench.py --no-cpu-baseline > gpurun_out/r17_bench_$i.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r17_bench_$i.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"; done
