#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/model_check.log) 2>&1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
for m in f16 tf32; do echo -n "$m: "; MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
timeout 300 python tools/time_cv.py 2>&1 | tail -2
