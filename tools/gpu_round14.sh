#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_convnet_gpu.py -m gpu -q > gpurun_out/pytest.log 2>&1; tail -12 gpurun_out/pytest.log | cut -c1-200
for h in 1 0; do MONOREC_B200_TC_HALO=$h MONOREC_B200_CONV=tf32 timeout 600 python tools/profile_model.py 8 4 3 2>&1 | tail -1; done
