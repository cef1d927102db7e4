#!/bin/bash
# round-2 first look: GPU suite, main bench line (library arm + nearest + C3D extras), C3D launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -15 gpurun_out/c1_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/c1_bench.json
timeout 300 python bench.py --network c3d --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c1_c3d_quick.json 2> gpurun_out/c1_c3d_quick.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/c1_c3d_launches.csv python bench.py --network c3d --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c1_c3d_ncu.log 2>&1
echo "ncu rc=$?"
