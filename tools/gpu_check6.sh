#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cost_volume_gpu.py -m gpu -q > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-full-model 2>/dev/null > gpurun_out/bench_pf0.json; python -c "import json; d=json.load(open('gpurun_out/bench_pf0.json')); print(round(d['value']), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],3), d['e2e']['value'])"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cost_volume -s 2 -c 1 python tools/profile_cv.py 2>&1 | grep -E "dram__|gpu__time"
