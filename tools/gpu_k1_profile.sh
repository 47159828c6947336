#!/bin/bash
# ncu captures of the cost-volume kernel at BASELINE config 2: full set with source, and the raw metrics as CSV
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cost_volume_kernel -s 2 -c 1 -f -o gpurun_out/prof_k1 python tools/profile_cv.py 8 4 4 > gpurun_out/ncu_k1.log 2>&1
tail -3 gpurun_out/ncu_k1.log
ncu -i gpurun_out/prof_k1.ncu-rep --page details > gpurun_out/prof_k1_details.txt 2>&1
ncu -i gpurun_out/prof_k1.ncu-rep --page raw --csv > gpurun_out/prof_k1_raw.csv 2>&1
ncu -i gpurun_out/prof_k1.ncu-rep --page source --csv > gpurun_out/prof_k1_source.csv 2>&1
ls -la gpurun_out/prof_k1*
