#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log | cut -c1-300
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $? stdout-lines $(wc -l < gpurun_out/bench.json)"
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read()); print(d['value'], d['e2e']['value'], d['full_model']['value'])"
