#!/bin/bash
# halo kernel with box rows of exactly 8 + kw - 1 pixels (more layers fit two CTAs per SM) against the fixed 16-px rows
mkdir -p gpurun_out
exec > >(tee gpurun_out/k2_pitch.log) 2>&1
timeout 900 python -m pytest tests/test_convnet_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python -m pytest tests/test_reprojection.py -q -m gpu 2>&1 | tail -2
for p in 0 16; do
  for m in f16 tf32; do
    echo "== pitch env $p, $m"
    MONOREC_B200_TC_HALO_PITCH=$p MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1
    MONOREC_B200_TC_HALO_PITCH=$p MONOREC_B200_CONV=$m timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -13
  done
done
