for i in 1 2; do
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_a$i.json 2> /dev/null
ZSV_DEBUG_WGRAD_MT1=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_b$i.json 2> /dev/null
done
