#!/bin/bash
mkdir -p gpurun_out
for h in 0 1; do MONOREC_B200_TC_HALO=$h timeout 600 python -m pytest tests/test_convnet_gpu.py -m gpu -q > gpurun_out/pytest_h$h.log 2>&1; echo "HALO=$h $(tail -1 gpurun_out/pytest_h$h.log)"; done
for h in 0 1; do echo "== HALO=$h"; MONOREC_B200_TC_HALO=$h timeout 300 python tools/bench_conv_layers.py 2>&1 | grep -v Downloading; done
for h in 0 1; do MONOREC_B200_TC_HALO=$h MONOREC_B200_CONV=tf32 timeout 600 python tools/profile_model.py 8 4 3 2>&1 | tail -1; done
