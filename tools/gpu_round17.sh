#!/bin/bash
mkdir -p gpurun_out
for T in 512 640 768; do
  MR_CV_THREADS=$T python -m monorec_b200.build --force > gpurun_out/build.log 2>&1
  timeout 300 python -m pytest tests/test_cost_volume_gpu.py -m gpu -x -q -k "golden" > gpurun_out/pytest_$T.log 2>&1
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-full-model > gpurun_out/bench_$T.json 2>/dev/null
  echo "thr=$T: $(tail -1 gpurun_out/pytest_$T.log) :: $(python -c "import json; d=json.load(open('gpurun_out/bench_$T.json')); print(round(d['value']), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],3))")"
done
