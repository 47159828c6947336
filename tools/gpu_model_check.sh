#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/model_check.log) 2>&1
timeout 900 python -m pytest tests/test_convnet_gpu.py -q -m gpu 2>&1 | tail -4
for m in f16 tf32; do echo -n "$m fused trunk: "; MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
for m in f16 tf32; do echo -n "$m separate ops: "; MONOREC_B200_TRUNK_FUSED=0 MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
