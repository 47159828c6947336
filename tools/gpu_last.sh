#!/bin/bash
# last call of the round: full suite under the default configuration + stage times
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/pytest_last.log 2>&1; tail -3 gpurun_out/pytest_last.log | cut -c1-300
for m in f16 tf32; do echo -n "default $m: "; MONOREC_B200_CONV=$m timeout 100 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
