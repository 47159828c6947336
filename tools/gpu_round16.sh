#!/bin/bash
for h in 0 1; do echo "== HALO=$h"; MONOREC_B200_TC_HALO=$h timeout 300 python tools/bench_conv_layers.py 2>&1 | grep -v Downloading; done
