#!/bin/bash
# validate the new linear / pool / Adam+pack kernels, then time both networks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -12 gpurun_out/c2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/c2_bench.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --optimizer torch > gpurun_out/c2_bench_torchadam.json 2> gpurun_out/c2_bench_torchadam.err; cut -c1-200 gpurun_out/c2_bench_torchadam.json
timeout 600 python bench.py --network c3d --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/c2_c3d.json 2> gpurun_out/c2_c3d.err; echo "c3d rc=$?"; cut -c1-300 gpurun_out/c2_c3d.json
timeout 300 python bench.py --network c3d --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c2_c3d_quick.json 2> gpurun_out/c2_c3d_quick.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/c2_c3d_launches.csv python bench.py --network c3d --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c2_c3d_ncu.log 2>&1
echo "ncu rc=$?"
