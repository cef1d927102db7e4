python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
SP="22,16,56,56,64,144,1,3,3,1,1,1,0,1,1"
S3="22,16,56,56,45,64,3,1,1,1,1,1,1,0,0"
python tools/bench_conv.py $SP $S3 > gpurun_out/wh_a.log 2>&1
ZSV_DEBUG_NO_WGRAD_HALO_SPATIAL=1 python tools/bench_conv.py $SP > gpurun_out/wh_b.log 2>&1
cat gpurun_out/wh_a.log gpurun_out/wh_b.log
for i in 1 2; do
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_a$i.json 2> /dev/null
ZSV_DEBUG_NO_WGRAD_HALO_SPATIAL=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_b$i.json 2> /dev/null
done
