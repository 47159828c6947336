#!/bin/bash
# K1 v2: compare 512- vs 256-thread builds (tests + bench + ncu each)
mkdir -p gpurun_out
for T in 512 256; do
  MR_CV_THREADS=$T python -m monorec_b200.build --force > gpurun_out/build_$T.log 2>&1
  timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$T.log
  timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume -s 2 -c 1 -o gpurun_out/prof_cv_v2_$T python tools/profile_cv.py > gpurun_out/ncu_full_$T.log 2>&1
  echo "== T=$T"; tail -4 gpurun_out/pytest_$T.log; cat gpurun_out/bench_$T.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('e2e',{}).get('value'), d['clocks'])"
done
