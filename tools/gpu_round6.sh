#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
grep -E "passed|failed|FAILED|Error|error|kitti|mask max|exit" gpurun_out/pytest.log | cut -c1-400 | tail -25
tail -4 gpurun_out/smoke.log
