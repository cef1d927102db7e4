#!/bin/bash
# Run every GPU test file in its own process (a faulting kernel poisons only its own CUDA context), each under
# a hard timeout, and collect logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
for f in "$@"; do
  name=$(basename "$f" .py)
  echo "=== $f ===" | tee -a gpurun_out/probe_summary.txt
  timeout 600 python -m pytest "$f" -m gpu -q --timeout 300 -p no:cacheprovider -s > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "rc=$rc" | tee -a gpurun_out/probe_summary.txt
  tail -n 25 "gpurun_out/$name.log" | tee -a gpurun_out/probe_summary.txt
done
