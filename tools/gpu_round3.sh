#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume -s 2 -c 1 -o gpurun_out/prof_cv_v3 python tools/profile_cv.py > gpurun_out/ncu_full.log 2>&1
grep -E "passed|failed|kitti|Error|error" gpurun_out/pytest.log | cut -c1-600 | tail -12
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('e2e',{}).get('value'), d['clocks'])"
