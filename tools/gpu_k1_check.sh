#!/bin/bash
# K1 bring-up: parity tests, sanitizer on small shapes, timing
mkdir -p gpurun_out
exec > >(tee gpurun_out/k1_check.log) 2>&1
timeout 600 python -m pytest tests/test_cost_volume_gpu.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/time_cv.py 2>&1 | tail -3
timeout 300 python tools/time_cv.py 4 6 64 512 1024 5 2>&1 | tail -3
bash tools/gpu_sanitize.sh 2>&1 | grep -v "^$" | tail -12
