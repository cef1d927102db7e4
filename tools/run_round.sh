python -m pytest tests/test_gpu_graph.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err
tail -4 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench11.err
