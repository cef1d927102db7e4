#!/bin/bash
# halo kernel with unrolled stage issue: conv parity tests, step time, layer table
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q > gpurun_out/c21_pytest.log 2>&1; tail -2 gpurun_out/c21_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-extras --layer-table > gpurun_out/c21_bench.json 2> gpurun_out/c21_layer_table.txt; cut -c1-300 gpurun_out/c21_bench.json
