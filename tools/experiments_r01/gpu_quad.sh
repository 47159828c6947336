#!/bin/bash
# epilogue store-pattern experiment: parity with quad stores, then stage times with/without, then the f16 launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_convnet_gpu.py -x -q > gpurun_out/pytest_quad.log 2>&1; tail -2 gpurun_out/pytest_quad.log | cut -c1-200
for q in 0 1; do for m in tf32 f16; do
  echo -n "quad=$q $m: "; MONOREC_B200_TC_QUAD=$q MONOREC_B200_CONV=$m timeout 300 python tools/profile_model.py 8 4 10 2>&1 | tail -1
done; done
MONOREC_B200_CONV=f16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/model_launches_f16.csv python tools/profile_model.py 8 4 1 > gpurun_out/ncu_model_f16.log 2>&1; echo "ncu exit $?"
