#!/bin/bash
# new fused BN-backward epilogue: conv tests first (each file under its own timeout), then A/B of the fuse modes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/c4_conv.log 2>&1; echo "conv rc=$?"
tail -5 gpurun_out/c4_conv.log
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_graph.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/c4_pytest.log
for mode in 2 1 0; do
  ZSV_FUSE_BN_BWD=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --layer-table > gpurun_out/c4_bench_fuse$mode.json 2> gpurun_out/c4_bench_fuse$mode.err; echo "bench fuse=$mode rc=$?"; cut -c1-200 gpurun_out/c4_bench_fuse$mode.json
done
timeout 300 python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c4_quick.json 2> gpurun_out/c4_quick.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c4_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c4_ncu.log 2>&1
echo "ncu rc=$?"
