#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/k2_ctas2.log) 2>&1
for c in 2 3 4; do
  echo "== halo CTA cap $c, f16"
  MONOREC_B200_TC_HALO=$c MONOREC_B200_CONV=f16 timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1
  MONOREC_B200_TC_HALO=$c MONOREC_B200_CONV=f16 timeout 200 python tools/bench_conv_layers.py 2>&1 | grep -E "enc0|dec3|dec4 3x3"
done
MONOREC_B200_CONV=f16 timeout 300 python tools/profile_layers.py 8 4 2>&1 | grep -v Warn | tail -42 | head -36
