#!/bin/bash
# 64-byte swizzle rows for 32-channel half chunks (parity first), halo variant with 2 CTAs/SM, per-layer timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_convnet_gpu.py -x -q -k "f16 or golden or half" > gpurun_out/pytest_k32.log 2>&1; tail -3 gpurun_out/pytest_k32.log | cut -c1-300
for k in 0 1; do echo "== f16 K32=$k"; MONOREC_B200_TC_K32=$k MONOREC_B200_CONV=f16 timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -7; done
for h in 0 1 2; do echo "== tf32 halo=$h"; MONOREC_B200_TC_HALO=$h MONOREC_B200_CONV=tf32 timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -7; done
echo -n "halo=2 parity: "; MONOREC_B200_TC_HALO=2 timeout 300 python -m pytest tests/test_convnet_gpu.py -x -q -k "tf32" 2>&1 | tail -1
for k in 0 1; do echo -n "f16 K32=$k: "; MONOREC_B200_TC_K32=$k MONOREC_B200_CONV=f16 timeout 300 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
