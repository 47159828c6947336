#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/last.log) 2>&1
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for z in 1 0; do echo -n "stem pool = $z, f16: "; MONOREC_B200_STEM_POOL=$z MONOREC_B200_CONV=f16 timeout 200 python tools/profile_model.py 8 4 20 2>&1 | tail -1; done
timeout 400 python bench.py --steps 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graph:', {k: d[k]['ms_per_forward'] for k in ('full_model','full_model_f16','full_model_f16_b16')}, d['value'])"
