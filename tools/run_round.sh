python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layer-table > gpurun_out/bench13.json 2> gpurun_out/bench13.err
tail -3 gpurun_out/pytest_gpu.log
