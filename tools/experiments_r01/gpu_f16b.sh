#!/bin/bash
mkdir -p gpurun_out
for m in tf32 f16; do MONOREC_B200_CONV=$m timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', d['full_model']['value'], d['full_model']['ms_per_forward'])"; done
MONOREC_B200_CONV=f16 timeout 300 python - <<'PY'
import sys, time, torch
sys.path.insert(0,'.')
from monorec_b200.model import MonoRecModel
from monorec_b200.synthetic import make_inputs, to_device
m = MonoRecModel().cuda().eval()
d = to_device(make_inputs(8,4,256,512,seed=0),'cuda:0')
with torch.no_grad():
    for _ in range(3): m(dict(d))
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(5): m(dict(d))
    t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
print(f"f16 eager: cpu issue {1e3*(t1-t0)/5:.2f} ms/forward, total {1e3*(t2-t0)/5:.2f} ms/forward")
PY
