#!/bin/bash
timeout 600 python tools/bench_torch_cuda.py 2>&1 | grep -v "Warning\|warn\|Downloading" | tail -6
