#!/bin/bash
# per-layer table of the forward (which convs are far from their HBM time) + the halo kernel forced to one CTA per SM
mkdir -p gpurun_out
exec > >(tee gpurun_out/k2_layers.log) 2>&1
for m in f16 tf32; do MONOREC_B200_CONV=$m timeout 300 python tools/profile_layers.py 8 4 2>&1 | grep -v Warn | tail -45; done
for m in f16 tf32; do echo "== halo forced, 1 CTA per SM, $m"; MONOREC_B200_TC_HALO=1 MONOREC_B200_CONV=$m timeout 200 python tools/bench_conv_layers.py 2>&1 | tail -13; done
