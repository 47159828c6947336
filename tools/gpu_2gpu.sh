#!/bin/bash
# 2-GPU checks: multi-rank parity test, bench under torchrun (both arms)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/pytest_2gpu.log 2>&1; tail -3 gpurun_out/pytest_2gpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "exit $? lines $(wc -l < gpurun_out/bench_2gpu.json)"
python -c "
import json; d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e']['value'], d['full_model']['value'])"
