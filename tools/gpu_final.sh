#!/bin/bash
# final verification of the round: full GPU suite, smoke, default + hi-res bench, reference arm, evidence captures
# (K1 is unchanged since its ncu captures of this round: profiles/r02_k1_*; the conv engine is re-captured)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $? lines $(wc -l < gpurun_out/bench.json)"
timeout 900 python bench.py --config hires --steps 20 --no-full-model --no-cpu-baseline > gpurun_out/bench_hires.json 2> gpurun_out/bench_hires.err; echo "hires exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-160 gpurun_out/bench_ref.json
for m in tf32 f16; do echo -n "$m: "; MONOREC_B200_CONV=$m timeout 300 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
MONOREC_B200_CONV=f16 timeout 300 python tools/profile_layers.py 8 4 2>&1 | grep -v Warn | tail -42 > gpurun_out/layers_f16.txt; head -3 gpurun_out/layers_f16.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 12 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-full-model > gpurun_out/ncu_bench.log 2>&1
MONOREC_B200_CONV=f16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo -s 2 -c 1 -f -o gpurun_out/prof_k2_f16 python tools/profile_conv.py > gpurun_out/ncu_k2.log 2>&1
ncu -i gpurun_out/prof_k2_f16.ncu-rep --page details > gpurun_out/prof_k2_f16_details.txt 2>&1
MONOREC_B200_CONV=f16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/model_launches_f16.csv python tools/profile_model.py 8 4 1 > gpurun_out/ncu_model_f16.log 2>&1
cat gpurun_out/bench.json | cut -c1-300
