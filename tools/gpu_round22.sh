#!/bin/bash
for c in 2 3 4; do echo "== CTAS=$c"; MONOREC_B200_TC_CTAS=$c timeout 300 python tools/bench_conv_layers.py 2>&1 | grep -v Downloading | head -4; done
