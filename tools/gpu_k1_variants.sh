#!/bin/bash
# K1: parity of the default build, then CUDA-event timings of every variant library under monorec_b200/variants/
mkdir -p gpurun_out
exec > >(tee gpurun_out/k1_variants.log) 2>&1
timeout 600 python -m pytest tests/test_cost_volume_gpu.py -q -m gpu 2>&1 | tail -6
echo "== default"; timeout 300 python tools/time_cv.py 2>&1 | tail -2
timeout 300 python tools/time_cv.py 4 6 64 512 1024 5 2>&1 | tail -2
for lib in monorec_b200/variants/*.so; do
  echo "== $(basename $lib)"
  MONOREC_B200_LIB=$PWD/$lib timeout 300 python tools/time_cv.py 2>&1 | tail -2 | head -1
  MONOREC_B200_LIB=$PWD/$lib timeout 300 python tools/time_cv.py 4 6 64 512 1024 5 2>&1 | tail -2 | head -1
done
for cfg in "4 6 64 512 1024"; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cost_volume_kernel -s 2 -c 1 --csv python tools/time_cv.py $cfg 3 2>/dev/null | grep -E "dram__bytes|gpu__time" | cut -d, -f 12-
done
bash tools/gpu_sanitize.sh 2>&1 | grep -v "^$" | grep "k1" | tail -4
