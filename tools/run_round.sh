python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_model.py tests/test_gpu_c3d.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
python tools/bench_bn.py > gpurun_out/bench_bn.log 2>&1
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench16.json 2> gpurun_out/bench16.err
cat gpurun_out/pytest_gpu.log gpurun_out/bench_bn.log
