#!/bin/bash
# re-take the fingerprinted DRAM-traffic / tensor-pipe capture of the conv kernels (part of tools/gpu_final.sh) after a csrc change
T=${1:-traffic}
mkdir -p gpurun_out
timeout 300 python bench.py --quick --no-graph --steps 1 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1 && \
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:igemm --csv --log-file gpurun_out/${T}_conv_traffic.csv python bench.py --quick --no-graph --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${T}_traffic_ncu.log 2>&1
python tools/conv_traffic.py gpurun_out/${T}_conv_traffic.csv profiles/r02_conv_dram_traffic.json | cut -c1-200
cp profiles/r02_conv_dram_traffic.json gpurun_out/${T}_conv_dram_traffic.json
