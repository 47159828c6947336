#!/bin/bash
# final verification of the round: full GPU suite, smoke, default + hi-res bench, reference arm, evidence captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $? lines $(wc -l < gpurun_out/bench.json)"
timeout 900 python bench.py --config hires --steps 20 --no-full-model --no-cpu-baseline > gpurun_out/bench_hires.json 2> gpurun_out/bench_hires.err; echo "hires exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-160 gpurun_out/bench_ref.json
for m in tf32 f16; do echo -n "$m: "; MONOREC_B200_CONV=$m timeout 300 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
MONOREC_B200_CONV=f16 timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -7
for cfg in "8 4 32 256 512" "4 6 64 512 1024"; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cost_volume_kernel -s 2 -c 1 --csv python tools/time_cv.py $cfg 3 2>/dev/null | grep -E "dram__bytes|gpu__time" | cut -d, -f 12-
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 12 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-full-model > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cost_volume_kernel -s 2 -c 1 -f -o gpurun_out/prof_k1_final python tools/profile_cv.py 8 4 4 > gpurun_out/ncu_k1.log 2>&1
ncu -i gpurun_out/prof_k1_final.ncu-rep --page details > gpurun_out/prof_k1_final_details.txt 2>&1
ncu -i gpurun_out/prof_k1_final.ncu-rep --page source --csv > gpurun_out/prof_k1_final_source.csv 2>&1
MONOREC_B200_CONV=f16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo -s 2 -c 1 -f -o gpurun_out/prof_k2_f16 python tools/profile_conv.py > gpurun_out/ncu_k2.log 2>&1
ncu -i gpurun_out/prof_k2_f16.ncu-rep --page details > gpurun_out/prof_k2_f16_details.txt 2>&1
MONOREC_B200_CONV=f16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/model_launches_f16.csv python tools/profile_model.py 8 4 1 > gpurun_out/ncu_model_f16.log 2>&1
cat gpurun_out/bench.json | cut -c1-300
