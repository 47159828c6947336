#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
timeout 600 python tools/profile_model.py 8 4 3 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/model_launches.csv python tools/profile_model.py 2 2 1 > gpurun_out/ncu_model.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/model_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.Counter(); cnt=collections.Counter()
for r in rows:
    name=r[4].split('(')[0][-60:]; agg[name]+=float(r[-1]); cnt[name]+=1
tot=sum(agg.values())
for n,t in agg.most_common(14): print(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}% x{cnt[n]:4d} {n}")
PY
