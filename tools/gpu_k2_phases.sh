#!/bin/bash
# sub-pixel phases in one launch + single-channel head epilogue: parity, then A/B of the one-launch switch
mkdir -p gpurun_out
exec > >(tee gpurun_out/k2_phases.log) 2>&1
timeout 900 python -m pytest tests/test_convnet_gpu.py -q -m gpu -x 2>&1 | tail -4
for o in 1 0; do
  for m in f16 tf32; do
    echo "== one launch $o, $m"
    MONOREC_B200_SUBPIXEL_ONE_LAUNCH=$o MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1
    MONOREC_B200_SUBPIXEL_ONE_LAUNCH=$o MONOREC_B200_CONV=$m timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -6
  done
done
MONOREC_B200_CONV=f16 timeout 300 python tools/profile_layers.py 8 4 2>&1 | grep -v Warn | tail -42 | head -24
