set -x
python -m pytest tests/test_gpu_graph.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_graph.log
python bench.py --steps 20 --warmup 5 --layer-table > gpurun_out/bench2.json 2> gpurun_out/bench2.err
tail -5 gpurun_out/pytest_graph.log; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
