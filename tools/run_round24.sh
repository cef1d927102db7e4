#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_model.py tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/r24_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r24_pytest.log
tail -4 gpurun_out/r24_pytest.log
for m in head new head new; do
  if [ $m = new ]; then unset ZSV_LIB_PATH; else export ZSV_LIB_PATH=build/ab/libzsv_head.so; fi
  python bench.py --no-cpu-baseline > gpurun_out/r24_bench_$m.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/r24_bench_$m.json').read().strip().splitlines()[-1]); print('$m', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
