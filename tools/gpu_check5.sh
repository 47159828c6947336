#!/bin/bash
for k in 0 2 3 4; do
  MR_CV_PF_ROWS=$k python -m monorec_b200.build --force > /dev/null 2>&1
  echo "PF_ROWS=$k $(timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-full-model 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['roofline']['kernel_ms'],3))")"
done
MR_CV_PF_ROWS=0 MR_CV_SKIP=3 python -m monorec_b200.build --force > /dev/null 2>&1
echo "PF_ROWS=0 stage1-only $(timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-full-model 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['roofline']['kernel_ms'],3))")"
python -m monorec_b200.build --force > /dev/null 2>&1
