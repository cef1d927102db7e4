#!/bin/bash
# final-state refresh: GPU suite, smoke, bench (+CPU arm), layer table, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/refresh_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/refresh_pytest.log
tail -3 gpurun_out/refresh_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/refresh_smoke.log 2>&1; tail -1 gpurun_out/refresh_smoke.log
python bench.py > gpurun_out/refresh_bench.json 2> gpurun_out/refresh_bench.err; cut -c1-200 gpurun_out/refresh_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/refresh_bench_reference.json 2> gpurun_out/refresh_bench_reference.err; cut -c1-300 gpurun_out/refresh_bench_reference.json
python bench.py --no-cpu-baseline --layer-table > /dev/null 2> gpurun_out/refresh_layer_table.txt
python bench.py --quick --no-graph --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/refresh_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/refresh_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/refresh_launches.csv > gpurun_out/refresh_launch_summary.txt; head -16 gpurun_out/refresh_launch_summary.txt
