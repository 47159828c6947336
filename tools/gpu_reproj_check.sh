#!/bin/bash
# reprojection loss: parity tests, memcheck of the same tests, timing at BASELINE config 2's image size
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reprojection.py -m gpu -q -s 2>&1 | tail -25 | cut -c1-220 | tee gpurun_out/reproj_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_reprojection.py -m gpu -q -x -k "golden or c_abi" > gpurun_out/reproj_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/reproj_memcheck.log | cut -c1-200
timeout 300 python tools/time_reprojection.py 2>&1 | tail -6 | tee gpurun_out/reproj_time.log
