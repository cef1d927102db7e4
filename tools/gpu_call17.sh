#!/bin/bash
mkdir -p gpurun_out
for p in 0 1 0 1; do ZSV_PDL=$p timeout 120 python tools/pdl_probe.py; done 2>&1 | tee gpurun_out/c17_pdl_probe.txt
