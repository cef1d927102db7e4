#!/bin/bash
mkdir -p gpurun_out
V=zeroshotvideoclassification_b200/build/variants
for v in cur2 early cur2 early; do
  echo "== $v"
  ZSV_LIB_PATH=$V/$v.so timeout 200 python tools/pdl_probe.py
  ZSV_LIB_PATH=$V/$v.so timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
ZSV_LIB_PATH=$V/early.so timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_graph.py -x -q 2>&1 | tail -2
