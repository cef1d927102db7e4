#!/bin/bash
# state refresh of the current tree: GPU suite, bench (all extras), layer table, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c20_pytest.log
tail -3 gpurun_out/c20_pytest.log
timeout 900 python bench.py > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err; cut -c1-200 gpurun_out/c20_bench.json
timeout 600 python bench.py --no-cpu-baseline --no-extras --layer-table > /dev/null 2> gpurun_out/c20_layer_table.txt
timeout 300 python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c20_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c20_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/c20_launches.csv > gpurun_out/c20_launch_summary.txt; head -30 gpurun_out/c20_launch_summary.txt
