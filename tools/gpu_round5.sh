#!/bin/bash
mkdir -p gpurun_out
for cfg in "512 1" "256 2" "256 1" "384 1"; do
  set -- $cfg
  MR_CV_THREADS=$1 MR_CV_MINBLOCKS=$2 python -m monorec_b200.build --force > gpurun_out/build.log 2>&1
  timeout 300 python -m pytest tests/test_cost_volume_gpu.py -m gpu -x -q -s > gpurun_out/pytest_$1_$2.log 2>&1
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  echo "== thr=$1 minblk=$2: $(tail -1 gpurun_out/pytest_$1_$2.log) :: $(python -c "import json; d=json.load(open('gpurun_out/bench_$1_$2.json')); print(round(d['value']), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],3))")"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume -s 2 -c 1 -o gpurun_out/prof_cv_v4_$1_$2 python tools/profile_cv.py > gpurun_out/ncu_full.log 2>&1
done
grep -h "kitti\|Error\|error" gpurun_out/pytest_512_1.log | cut -c1-700 | tail -5
