#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json; cut -c1-300 gpurun_out/bench_ref.json
