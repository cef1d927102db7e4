#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c26_pytest.log 2>&1; tail -2 gpurun_out/c26_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-extras --layer-table > gpurun_out/c26_bench.json 2> gpurun_out/c26_layer_table.txt; cut -c1-300 gpurun_out/c26_bench.json
