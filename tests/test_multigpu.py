"""1-vs-N-rank equality of the gathered result (needs >= 2 GPUs; skipped on a 1-GPU box)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MR_ROOT"])
from monorec_b200.dist import all_gather_batch, shard_data_dict
from monorec_b200.model import MonoRecModel
from monorec_b200.synthetic import make_inputs, seeded_state_dict, to_device
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
model = MonoRecModel()
model.load_state_dict(seeded_state_dict(model, seed=7, gain=0.7))
model = model.to(dev).eval()
data = make_inputs(4, 2, 64, 128, seed=17)                     # the same global batch on every rank
full = model(to_device(data, dev))["result"]                   # unsharded reference on this rank
mine = model(to_device(shard_data_dict(data, rank, world), dev))["result"]
gathered = all_gather_batch(mine)
ok = torch.equal(gathered, full)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_OK" if int(flag.item()) == 1 else "MULTIGPU_MISMATCH", tuple(gathered.shape))
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_forward_all_gather_equals_single_rank(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MR_ROOT=str(ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert "MULTIGPU_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])
