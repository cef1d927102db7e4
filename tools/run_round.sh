set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --layer-table --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --quick --no-graph --steps 1 --warmup 1 > gpurun_out/ncu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench4.json
