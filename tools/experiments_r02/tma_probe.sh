#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tools/experiments_r02/tma_probe.cu || exit 1
for mode in 0 1; do for rank in 2 3; do for bw in 32 64 128; do
  timeout 60 /tmp/tma_probe $rank $bw 8 16 8 $mode
done; done; done 2>&1 | tee gpurun_out/tma_probe.txt
for cfg in "3 128 8 -2 -2 1" "3 128 8 300 44 1" "3 128 8 -2 16 1" "3 64 8 -2 -2 1" "2 128 8 -2 8 1" "3 128 8 400 60 1"; do timeout 60 /tmp/tma_probe $cfg; done 2>&1 | tee -a gpurun_out/tma_probe.txt
