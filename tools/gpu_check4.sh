#!/bin/bash
for k in 0 1 2 3 4; do
  MR_CV_SKIP=$k python -m monorec_b200.build --force > /dev/null 2>&1
  echo "SKIP=$k $(timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-full-model 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['roofline']['kernel_ms'],3))")"
done
python -m monorec_b200.build --force > /dev/null 2>&1
