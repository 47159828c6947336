#!/bin/bash
# First GPU call of the next round: measure what round 1 could only prepare (no GPU minutes were left).
#   1. tools/ubench_tc.cu: tcgen05.mma issue cost vs shape / chains / CTAs, TMA box rates for the boxes the conv kernels use
#   2. parity + timing of the opt-in conv variants: staged epilogue (MONOREC_B200_TC_EPI=1), 64-byte rows inside the halo box
#      (MONOREC_B200_TC_HALO_K32=1), 8 epilogue warps over 4 accumulators in the single-CTA halo kernel
#      (MONOREC_B200_TC_HALO_EPI8=1), and their combinations
mkdir -p gpurun_out
exec > >(tee gpurun_out/next_round.log) 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_tc tools/ubench_tc.cu && timeout 300 /tmp/ubench_tc > gpurun_out/ubench_tc.txt 2>&1; tail -5 gpurun_out/ubench_tc.txt
run() {  # name, env assignments...
  local name=$1; shift
  echo "== $name"
  env "$@" timeout 600 python -m pytest tests/test_convnet_gpu.py -x -q 2>&1 | tail -4
  for m in f16 tf32; do echo -n "$m: "; env "$@" MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
  env "$@" MONOREC_B200_CONV=f16 timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -7
}
run "default" MONOREC_B200_NOOP=1
run "staged epilogue" MONOREC_B200_TC_EPI=1
run "staged epilogue, no halo" MONOREC_B200_TC_EPI=1 MONOREC_B200_TC_HALO=0
run "halo with 64-byte rows" MONOREC_B200_TC_HALO_K32=1
run "staged epilogue + 64-byte halo rows" MONOREC_B200_TC_EPI=1 MONOREC_B200_TC_HALO_K32=1
run "staged epilogue + 64-byte halo rows + 8-warp halo epilogue" MONOREC_B200_TC_EPI=1 MONOREC_B200_TC_HALO_K32=1 MONOREC_B200_TC_HALO_EPI8=1
run "single-channel heads on the tensor cores" MONOREC_B200_TC_HEADS=1
run "cuDNN trunk in half (half mode only)" MONOREC_B200_TRUNK=cudnn_f16
echo "== staged goldens on the DEFAULT path: bundled-sample full model, 64 planes x 6 frames"
MONOREC_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_experimental_gpu.py -q -m gpu -s -k "bundled or d64 or maxpool" 2>&1 | tail -12
echo "== experimental pieces: 3x3/s2 max-pool, residual epilogue, ResNet trunk on the engine"
MONOREC_B200_EXPERIMENTAL=1 MONOREC_B200_TC_EPI=1 timeout 600 python -m pytest tests/test_experimental_gpu.py -q -m gpu 2>&1 | tail -3
for m in f16 tf32; do echo -n "engine trunk $m: "; MONOREC_B200_TC_EPI=1 MONOREC_B200_TRUNK=engine MONOREC_B200_CONV=$m timeout 200 python tools/profile_model.py 8 4 10 2>&1 | tail -1; done
MONOREC_B200_TC_EPI=1 MONOREC_B200_TRUNK=engine timeout 600 python -m pytest tests/test_convnet_gpu.py -q -m gpu -k "golden or graph" 2>&1 | tail -1
