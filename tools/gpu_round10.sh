#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_convnet_gpu.py -m gpu -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
MONOREC_B200_CONV=tf32 timeout 600 python tools/profile_model.py 8 4 3 2>&1 | tail -1
MONOREC_B200_CONV=tf32 timeout 600 python - <<'PY'
import sys, time, torch
sys.path.insert(0,'.')
from monorec_b200.model import MonoRecModel, GraphedMonoRec
from monorec_b200.synthetic import make_inputs, to_device
for B in (8, 16):
    m = MonoRecModel().cuda().eval()
    d = to_device(make_inputs(B,4,256,512,seed=0),'cuda:0')
    g = GraphedMonoRec(m, d)
    ref = m(dict(d))["result"].clone()
    out = g(d)
    torch.cuda.synchronize()
    print("graph vs eager max|d|", float((out["result"]-ref).abs().max()))
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    for _ in range(3): g(d)
    e0.record()
    for _ in range(10): g(d)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(f"graphed B={B}: {ms:.2f} ms/forward = {1e3*B/ms:.1f} keyframes/s")
PY
MONOREC_B200_CONV=tf32 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/model_launches_tf32.csv python tools/profile_model.py 8 4 1 > gpurun_out/ncu_model.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/model_launches_tf32.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.Counter(); cnt=collections.Counter()
for r in rows:
    name=r[4].split('(')[0][-50:]; agg[name]+=float(r[-1]); cnt[name]+=1
tot=sum(agg.values())
for n,t in agg.most_common(5): print(f"{t/7e6:9.3f} ms/fwd {100*t/tot:5.1f}% x{cnt[n]//7:4d} {n}")
PY
