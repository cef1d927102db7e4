python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -3 gpurun_out/bench_8gpu.err; cat gpurun_out/bench_8gpu.json
