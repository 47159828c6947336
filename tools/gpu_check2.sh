#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_convnet_gpu.py -m gpu -q > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
MONOREC_B200_CONV=tf32 timeout 600 python tools/profile_model.py 8 4 3 2>&1 | tail -1
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['full_model'])"
