#!/bin/bash
# bench.py on one GPU (default and hires), reference arm, ncu DRAM traffic + full capture of the final cost-volume kernel
mkdir -p gpurun_out
exec > >(tee gpurun_out/bench_check.log) 2>&1
timeout 900 python bench.py --steps 50 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-3000
timeout 900 python bench.py --config hires --steps 20 --no-full-model --no-cpu-baseline > gpurun_out/bench_hires.json 2> gpurun_out/bench_hires.err; echo "hires exit $?"; tail -3 gpurun_out/bench_hires.err; cat gpurun_out/bench_hires.json | cut -c1-2000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-400 gpurun_out/bench_ref.json
for cfg in "8 4 32 256 512" "4 6 64 512 1024"; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cost_volume_kernel -s 2 -c 1 --csv python tools/time_cv.py $cfg 3 2>/dev/null | grep -E "dram__bytes|gpu__time" | cut -d, -f 12-
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cost_volume_kernel -s 2 -c 1 -f -o gpurun_out/prof_k1_final python tools/profile_cv.py 8 4 4 > gpurun_out/ncu_k1.log 2>&1
ncu -i gpurun_out/prof_k1_final.ncu-rep --page details > gpurun_out/prof_k1_final_details.txt 2>&1
ncu -i gpurun_out/prof_k1_final.ncu-rep --page source --csv > gpurun_out/prof_k1_final_source.csv 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 40 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-full-model > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/ | tail -12
