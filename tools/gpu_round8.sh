#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_convnet_gpu.py -m gpu -q -s -x -k "tc_conv" > gpurun_out/pytest_tc1.log 2>&1; echo "exit $?" >> gpurun_out/pytest_tc1.log
tail -15 gpurun_out/pytest_tc1.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "passed|failed|FAILED|tf32|fp32 g|exit" gpurun_out/pytest.log | cut -c1-300 | tail -30
for m in fp32 tf32; do MONOREC_B200_CONV=$m timeout 600 python tools/profile_model.py 8 4 3 2>&1 | tail -1; done
