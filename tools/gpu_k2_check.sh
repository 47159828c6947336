#!/bin/bash
# K2 after the clean-up (staged epilogue + 64-byte halo rows by default): parity suite with the numbers printed, stage times
mkdir -p gpurun_out
exec > >(tee gpurun_out/k2_check.log) 2>&1
timeout 900 python -m pytest tests/test_convnet_gpu.py -q -m gpu -s 2>&1 | grep -E "bundled sample|mask max|passed|failed|FAILED|Error" | head -60
for m in f16 tf32; do echo -n "$m: "; MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
for m in f16 tf32; do echo -n "tc heads $m: "; MONOREC_B200_TC_HEADS=1 MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
MONOREC_B200_TC_HEADS=1 timeout 900 python -m pytest tests/test_convnet_gpu.py -q -m gpu 2>&1 | tail -3
MONOREC_B200_CONV=f16 timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -7
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3
