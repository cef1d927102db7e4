#!/bin/bash
# full-state refresh: GPU suite, smoke, bench, launch list, layer table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r16_pytest.log
tail -3 gpurun_out/r16_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r16_smoke.log 2>&1; tail -1 gpurun_out/r16_smoke.log
python bench.py > gpurun_out/r16_bench.json 2> gpurun_out/r16_bench.err; cat gpurun_out/r16_bench.json | cut -c1-400
python bench.py --no-cpu-baseline --layer-table > /dev/null 2> gpurun_out/r16_layer_table.txt
python bench.py --no-graph --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r16_launches.csv python bench.py --no-graph --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r16_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r16_launches.csv > gpurun_out/r16_launch_summary.txt; head -12 gpurun_out/r16_launch_summary.txt
