#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_convnet_gpu.py -m gpu -q -s -k "f16" > gpurun_out/pytest_f16.log 2>&1; grep -E "passed|failed|FAILED|f16 g07|Error" gpurun_out/pytest_f16.log | cut -c1-250 | tail -12
timeout 900 python -m pytest tests/test_convnet_gpu.py -m gpu -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log | cut -c1-200
for m in tf32 f16; do MONOREC_B200_CONV=$m timeout 600 python tools/profile_model.py 8 4 3 2>&1 | tail -1; done
