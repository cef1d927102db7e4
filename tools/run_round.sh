timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for i in 1 2; do
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_a$i.json 2> /dev/null
ZSV_DEBUG_KEEP_NSTG2=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_b$i.json 2> /dev/null
ZSV_2CTA=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_c$i.json 2> /dev/null
done
