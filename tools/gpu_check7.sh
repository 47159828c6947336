#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cost_volume_gpu.py -m gpu -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log | cut -c1-300
for pk in 1 0; do echo "PACKED=$pk $(MONOREC_B200_CV_PACKED=$pk timeout 300 python tools/profile_model.py 8 4 5 2>&1 | tail -1)"; done
