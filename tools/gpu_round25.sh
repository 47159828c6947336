#!/bin/bash
mkdir -p gpurun_out
for ws in 0 1; do
  MR_CV_WS=$ws python -m monorec_b200.build --force > gpurun_out/build.log 2>&1
  timeout 240 python -m pytest tests/test_cost_volume_gpu.py -m gpu -x -q > gpurun_out/pytest_ws$ws.log 2>&1
  timeout 240 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-full-model > gpurun_out/bench_ws$ws.json 2>/dev/null
  echo "WS=$ws: $(tail -1 gpurun_out/pytest_ws$ws.log) :: $(python -c "import json; d=json.loads(open('gpurun_out/bench_ws$ws.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],3))")"
done
