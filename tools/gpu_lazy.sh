#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/lazy.log) 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
for z in 1 0; do for m in f16 tf32; do echo -n "lazy level 4 = $z, $m: "; MONOREC_B200_TRUNK_LAZY_LEVEL4=$z MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 20 2>&1 | tail -1; done; done
timeout 600 python bench.py --steps 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graph:', {k: d[k]['ms_per_forward'] for k in ('full_model','full_model_f16','full_model_f16_b16')})"
