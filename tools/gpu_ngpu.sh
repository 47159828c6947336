#!/bin/bash
# N-GPU checks (N = $1, default 4): multi-rank parity test, bench under torchrun exactly as the driver launches it (both arms)
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/pytest_ngpu.log 2>&1; tail -3 gpurun_out/pytest_ngpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "exit $? lines $(wc -l < gpurun_out/bench_${N}gpu.json)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${N}gpu_ref.json 2>/dev/null; echo "ref exit $?"; cut -c1-120 gpurun_out/bench_${N}gpu_ref.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['full_model']['value'], d['full_model_f16']['value'], d['config'])"
