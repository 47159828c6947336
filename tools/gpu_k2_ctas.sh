#!/bin/bash
# halo kernel: up to 4 CTAs per SM (automatic) against the cap of 2
mkdir -p gpurun_out
exec > >(tee gpurun_out/k2_ctas.log) 2>&1
timeout 900 python -m pytest tests/test_convnet_gpu.py -q -m gpu -x 2>&1 | tail -2
for c in 4 3 2; do
  for m in f16 tf32; do
    echo "== halo CTA cap $c, $m"
    MONOREC_B200_TC_HALO=$c MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1
    MONOREC_B200_TC_HALO=$c MONOREC_B200_CONV=$m timeout 200 python tools/bench_conv_layers.py 2>&1 | grep -v refine | grep -v upconv | tail -13
  done
done
