#!/bin/bash
# pair wgrad kernel: conv tests, then A/B of the training step with and without it (layer table of each)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/c7_conv.log 2>&1; echo "conv rc=$?"
tail -5 gpurun_out/c7_conv.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --layer-table > gpurun_out/c7_bench_pair.json 2> gpurun_out/c7_lt_pair.txt; echo "rc=$?"; cut -c1-200 gpurun_out/c7_bench_pair.json
ZSV_WGRAD_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --layer-table > gpurun_out/c7_bench_nopair.json 2> gpurun_out/c7_lt_nopair.txt; echo "rc=$?"; cut -c1-200 gpurun_out/c7_bench_nopair.json
ZSV_FUSE_BN_BWD=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/c7_bench_fuse1.json 2> /dev/null; cut -c1-200 gpurun_out/c7_bench_fuse1.json
