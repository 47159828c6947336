#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['full_model']['value'], d['full_model']['ms_per_forward'], d['cpu_baseline']['value'], d['clocks'])"
