#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'])"; }
for rep in 1 2; do
ZSV_OVERLAP_WGRAD=0 ZSV_HALO_WSHIFT=0 run "overlap0 wshift0"
ZSV_OVERLAP_WGRAD=0 run "overlap0 wshift1"
ZSV_HALO_WSHIFT=0 run "overlap1 wshift0"
run "overlap1 wshift1"
ZSV_HALO_2CTA=0 ZSV_2CTA=0 run "overlap1 nopairs"
done
