run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2> /dev/null; }
run 29511 > gpurun_out/n2_default.json
NCCL_MIN_CTAS=16 run 29512 > gpurun_out/n2_min16.json
NCCL_MIN_CTAS=32 run 29513 > gpurun_out/n2_min32.json
ZSV_BUCKET_MB=8 run 29514 > gpurun_out/n2_b8.json
ZSV_BUCKET_MB=200 run 29515 > gpurun_out/n2_b200.json
