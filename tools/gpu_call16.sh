#!/bin/bash
# programmatic dependent launch: full GPU suite, then A/B of the training step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/c16_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/c16_pytest.log
for pdl in 1 0 1 0; do
ZSV_PDL=$pdl timeout 600 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=$pdl', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done 2>&1 | tee gpurun_out/c16_pdl_ab.txt
ZSV_PDL=1 timeout 600 python bench.py --network c3d --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | cut -c1-200 | tee -a gpurun_out/c16_pdl_ab.txt
ZSV_PDL=0 timeout 600 python bench.py --network c3d --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | cut -c1-200 | tee -a gpurun_out/c16_pdl_ab.txt
