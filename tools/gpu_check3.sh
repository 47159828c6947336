#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cost_volume_gpu.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-full-model 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],3))"
