set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --layer-table > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --quick --no-graph --steps 1 --warmup 1 > gpurun_out/plain_quick.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --quick --no-graph --steps 1 --warmup 1 > gpurun_out/ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:igemm" -c 600 --csv --log-file gpurun_out/conv_traffic.csv python bench.py --quick --no-graph --steps 1 --warmup 1 > gpurun_out/ncu2.log 2>&1
python tools/bench_bn.py > gpurun_out/bench_bn.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
tail -3 gpurun_out/pytest_gpu.log
