#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/r19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r19_pytest.log
tail -5 gpurun_out/r19_pytest.log
for m in old new old new; do
  if [ $m = new ]; then unset ZSV_DEBUG_PACK_ELEMENTWISE; else export ZSV_DEBUG_PACK_ELEMENTWISE=1; fi
  python bench.py --no-cpu-baseline > gpurun_out/r19_bench_$m.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/r19_bench_$m.json').read().strip().splitlines()[-1]); print('$m', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
