#!/bin/bash
# last call of the round: full suite under the default configuration, then the half-precision halo experiment
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_last.log 2>&1; tail -3 gpurun_out/pytest_last.log | cut -c1-300
echo -n "f16 halo parity: "; MONOREC_B200_TC_HALO_F16=1 timeout 300 python -m pytest tests/test_convnet_gpu.py -x -q -k "f16" 2>&1 | tail -1
for m in tf32 f16; do echo -n "default $m: "; MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
echo -n "halo_f16 f16: "; MONOREC_B200_TC_HALO_F16=1 MONOREC_B200_CONV=f16 timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1
echo "== f16 default"; MONOREC_B200_CONV=f16 timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -7
echo "== f16 halo_f16"; MONOREC_B200_TC_HALO_F16=1 MONOREC_B200_CONV=f16 timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -7
echo "== tf32 default"; MONOREC_B200_CONV=tf32 timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -7
