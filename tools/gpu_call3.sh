#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
tail -8 gpurun_out/c3_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --layer-table > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/c3_bench.json
timeout 300 python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c3_quick.json 2> gpurun_out/c3_quick.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c3_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c3_ncu.log 2>&1
echo "ncu rc=$?"
