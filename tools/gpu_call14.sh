#!/bin/bash
# 2-GPU data-parallel step: scaling and replica consistency with the gradient arena
mkdir -p gpurun_out
NCCL_DEBUG=VERSION timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/c14_bench_2gpu.json 2> gpurun_out/c14_bench_2gpu.err; echo "rc=$?"
cut -c1-250 gpurun_out/c14_bench_2gpu.json; grep -o '"replicas[^}]*}' gpurun_out/c14_bench_2gpu.json | head -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/c14_bench_1gpu.json 2> /dev/null; cut -c1-250 gpurun_out/c14_bench_1gpu.json
