#!/bin/bash
run() { echo "$1 :: $(timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-full-model 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],3))")"; }
MR_CV_TILE_ROWS=32 python -m monorec_b200.build --force > /dev/null 2>&1; run "TH=32"
MR_CV_THREADS=256 MR_CV_MINBLOCKS=2 python -m monorec_b200.build --force > /dev/null 2>&1; run "256x2"
MR_CV_THREADS=256 MR_CV_MINBLOCKS=1 python -m monorec_b200.build --force > /dev/null 2>&1; run "256x1"
python -m monorec_b200.build --force > /dev/null 2>&1; run "default"
