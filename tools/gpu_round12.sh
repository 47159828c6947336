#!/bin/bash
mkdir -p gpurun_out
for TH in 16 32; do
  MR_CV_TILE_ROWS=$TH python -m monorec_b200.build --force > gpurun_out/build.log 2>&1
  timeout 300 python -m pytest tests/test_cost_volume_gpu.py -m gpu -x -q > gpurun_out/pytest_th$TH.log 2>&1
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-full-model > gpurun_out/bench_th$TH.json 2>/dev/null
  echo "TH=$TH: $(tail -1 gpurun_out/pytest_th$TH.log) :: $(python -c "import json; d=json.load(open('gpurun_out/bench_th$TH.json')); print(round(d['value']), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],3))")"
done
python -m monorec_b200.build --force > gpurun_out/build.log 2>&1
# evidence for profiles/: launch list of the bench command, full captures of K1 and of the tensor-core conv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 40 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-full-model > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume -s 2 -c 1 -o gpurun_out/prof_k1 python tools/profile_cv.py > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof_k2_fullres python tools/profile_conv.py > gpurun_out/ncu_k2a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 7 -c 1 -o gpurun_out/prof_k2_refine python tools/profile_conv.py > gpurun_out/ncu_k2b.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
