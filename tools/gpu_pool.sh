#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/pool.log) 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
for z in 1 0; do for m in f16 tf32; do echo -n "fused pool = $z, $m: "; MONOREC_B200_FUSED_POOL=$z MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 20 2>&1 | tail -1; done; done
