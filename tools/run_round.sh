set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --layer-table > gpurun_out/bench1.json 2> gpurun_out/bench1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 1 --warmup 1 > gpurun_out/ncu.log 2>&1
python tools/bench_bn.py > gpurun_out/bench_bn.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench1.json
