#!/bin/bash
# last call of the round: the driver's three commands on the final tree
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log | cut -c1-150
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], {k: d[k]['ms_per_forward'] for k in ('full_model','full_model_f16','full_model_f16_b16')}, d.get('reprojection_loss',{}).get('forward_backward_ms'), [k for k in d if 'error' in k])"
