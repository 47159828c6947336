#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $? lines $(wc -l < gpurun_out/bench.json)"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], d['e2e']['value'], d['full_model']['value'], d['cpu_baseline']['value'], d['clocks'])"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"cost_volume|repack" -s 4 -c 2 python tools/profile_cv.py 2>&1 | grep -E "dram__|gpu__time|repack|cost_volume_kernel"
