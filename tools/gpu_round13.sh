#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_convnet_gpu.py -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
MONOREC_B200_CONV=tf32 timeout 600 python tools/profile_model.py 8 4 3 2>&1 | tail -1
MONOREC_B200_CONV=tf32 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/model_launches_tf32.csv python tools/profile_model.py 8 4 1 > gpurun_out/ncu_model.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/model_launches_tf32.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.Counter(); cnt=collections.Counter()
for r in rows:
    name=r[4].split('(')[0][-50:]; agg[name]+=float(r[-1]); cnt[name]+=1
tot=sum(agg.values())
for n,t in agg.most_common(6): print(f"{t/7e6:9.3f} ms/fwd {100*t/tot:5.1f}% x{cnt[n]//7:4d} {n}")
tc=sorted([float(r[-1])/1e3 for r in rows[-len(rows)//7:] if 'conv_tc' in r[4]], reverse=True)
print("top tc launches us:", [round(x) for x in tc[:12]])
PY
