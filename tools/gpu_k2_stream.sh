#!/bin/bash
# halo kernel with streamed weights (layers whose weights do not fit in shared memory) against the tap-refetch kernel
mkdir -p gpurun_out
exec > >(tee gpurun_out/k2_stream.log) 2>&1
timeout 900 python -m pytest tests/test_convnet_gpu.py -q -m gpu -x 2>&1 | tail -3
for o in 1 0; do
  for m in f16 tf32; do
    echo "== stream $o, $m"
    MONOREC_B200_TC_STREAM=$o MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1
    MONOREC_B200_TC_STREAM=$o MONOREC_B200_CONV=$m timeout 200 python tools/bench_conv_layers.py 2>&1 | grep -E "dec3|dec1|enc1 3x3|7x1|1x7"
  done
done
MONOREC_B200_CONV=f16 timeout 300 python tools/profile_layers.py 8 4 2>&1 | grep -v Warn | tail -42 | head -30
