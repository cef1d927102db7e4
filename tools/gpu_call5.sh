#!/bin/bash
# 2-GPU data-parallel run: gradient arena + hooks + replica-consistency check; 1-GPU line of the same build beside it
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --layer-table > gpurun_out/c5_bench1.json 2> gpurun_out/c5_bench1.err; echo "bench1 rc=$?"; cut -c1-200 gpurun_out/c5_bench1.json
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c5_bench2.json 2> gpurun_out/c5_bench2.err; echo "bench2 rc=$?"; cut -c1-300 gpurun_out/c5_bench2.json
grep -E "NVLS|nranks|Connected|channels" gpurun_out/c5_bench2.err | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --network c3d > gpurun_out/c5_c3d2.json 2> gpurun_out/c5_c3d2.err; echo "c3d2 rc=$?"; cut -c1-300 gpurun_out/c5_c3d2.json
