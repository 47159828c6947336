#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tools/experiments_r02/tma_probe.cu || exit 1
for cfg in "3 128 8 2 8 1" "3 128 8 3 8 1" "3 128 8 4 8 1" "3 128 8 -4 8 1" "3 128 8 -16 8 1" "3 128 8 0 -2 1" "3 128 8 -4 -2 1" "3 128 8 -128 8 1" "3 128 8 -132 -9 1" "3 128 8 330 47 1" "3 128 8 1 8 1" "3 32 8 -4 8 0" "2 32 8 -4 8 0"; do timeout 60 /tmp/tma_probe $cfg; done 2>&1 | tee gpurun_out/tma_probe2.txt
