#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cost_volume_gpu.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-full-model > gpurun_out/bench_hint.json 2>/dev/null
python -c "import json; d=json.loads(open('gpurun_out/bench_hint.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],3))"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:cost_volume -s 2 -c 1 python tools/profile_cv.py 2>&1 | grep -E "dram__|gpu__time|lts__"
