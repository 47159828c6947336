#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/k2_debug.log) 2>&1
python -m pytest tests/test_convnet_gpu.py -q -m gpu -k "bundled_sample" -s 2>&1 | grep -E "bundled sample|passed|failed" | cut -c1-250
echo "--- stream off"
MONOREC_B200_TC_STREAM=0 python -m pytest tests/test_convnet_gpu.py -q -m gpu -k "bundled_sample and f16" -s 2>&1 | grep -E "bundled sample|passed|failed" | cut -c1-250
timeout 900 python -m pytest tests/test_convnet_gpu.py -q -m gpu 2>&1 | tail -4
for m in f16 tf32; do echo -n "$m: "; MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
MONOREC_B200_CONV=f16 timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -19
