#!/bin/bash
mkdir -p gpurun_out
MONOREC_B200_CONV=tf32 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/model_launches_tf32.csv python tools/profile_model.py 8 4 1 > gpurun_out/ncu_model.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/model_launches_tf32.csv')) if len(r)>10 and r[0].isdigit()]
# last forward only: take the final third of launches roughly -> aggregate all and divide
agg=collections.Counter(); cnt=collections.Counter()
for r in rows:
    name=r[4].split('(')[0][-50:]; agg[name]+=float(r[-1]); cnt[name]+=1
tot=sum(agg.values())
for n,t in agg.most_common(8): print(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}% x{cnt[n]:4d} {n}")
# top individual conv_tc launches with grid sizes
tc=[r for r in rows if 'conv_tc' in r[4]]
tc.sort(key=lambda r:-float(r[-1]))
for r in tc[:12]: print(r[8], r[7], float(r[-1])/1e3, "us")
PY
MONOREC_B200_CONV=tf32 timeout 600 python - <<'PY'
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
    t1=time.perf_counter()   # CPU-side issue time (no sync)
    torch.cuda.synchronize()
    t2=time.perf_counter()
print(f"cpu issue {1e3*(t1-t0)/5:.2f} ms/forward, total {1e3*(t2-t0)/5:.2f} ms/forward")
PY
