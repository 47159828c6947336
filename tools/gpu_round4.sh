#!/bin/bash
mkdir -p gpurun_out
best=0
for cfg in "256 1 1" "256 2 0" "512 1 0" "256 2 1" "256 1 0"; do
  set -- $cfg
  MR_CV_THREADS=$1 MR_CV_MINBLOCKS=$2 MR_CV_PREFETCH=$3 python -m monorec_b200.build --force > gpurun_out/build.log 2>&1
  timeout 300 python -m pytest tests/test_cost_volume_gpu.py -m gpu -x -q -k "golden_small or golden_kitti" > gpurun_out/pytest_$1_$2_$3.log 2>&1
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$1_$2_$3.json 2> gpurun_out/bench_$1_$2_$3.err
  echo "== thr=$1 minblk=$2 prefetch=$3: $(tail -1 gpurun_out/pytest_$1_$2_$3.log) :: $(python -c "import json; d=json.load(open('gpurun_out/bench_$1_$2_$3.json')); print(round(d['value']), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],3))")"
done
# profile the default-variant again for the record
MR_CV_THREADS=256 MR_CV_MINBLOCKS=2 MR_CV_PREFETCH=0 python -m monorec_b200.build --force > gpurun_out/build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume -s 2 -c 1 -o gpurun_out/prof_cv_v3b python tools/profile_cv.py > gpurun_out/ncu_full.log 2>&1
