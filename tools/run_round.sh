for w in 2 1 3 2 1; do
ZSV_DEBUG_WGRAD_WAVES=$w python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_w$w.json 2> /dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/ab_w$w.json'))
print('waves', $w, round(d['value'],1), round(d['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['roofline']['by_kernel'].items()})
PY
done > gpurun_out/ab_waves.log 2>&1
cat gpurun_out/ab_waves.log
