#!/bin/bash
# round-2 re-entry baseline: GPU suite, main bench line with extras (library arm, nearest, C3D), layer table, launch lists
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
tail -8 gpurun_out/c6_pytest.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/c6_bench.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --layer-table > gpurun_out/c6_bench_lt.json 2> gpurun_out/c6_layer_table.txt; echo "lt rc=$?"
timeout 300 python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c6_quick.json 2> gpurun_out/c6_quick.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c6_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c6_ncu.log 2>&1
echo "ncu rc=$?"
timeout 300 python bench.py --network c3d --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c6_c3d_quick.json 2> gpurun_out/c6_c3d_quick.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/c6_c3d_launches.csv python bench.py --network c3d --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c6_c3d_ncu.log 2>&1
echo "c3d ncu rc=$?"
