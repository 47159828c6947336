#!/bin/bash
mkdir -p gpurun_out
MONOREC_B200_TC_HALO=1 timeout 600 python -m pytest tests/test_convnet_gpu.py -m gpu -q -k "tc_conv" > gpurun_out/pytest_halo.log 2>&1; tail -8 gpurun_out/pytest_halo.log | cut -c1-200
timeout 600 python -m pytest tests/test_convnet_gpu.py -m gpu -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
