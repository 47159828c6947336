"""N>1 host logic on CPU: world_size-2 gloo run of the shard / gather helpers."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from monorec_b200.dist import all_gather_batch, shard_bounds, shard_data_dict
    from monorec_b200.synthetic import make_inputs
    data = make_inputs(batch, 2, 16, 32, seed=3)
    mine = shard_data_dict(data, rank, world)
    lo, hi = shard_bounds(batch, rank, world)
    assert mine["keyframe"].shape[0] == hi - lo and mine["frames"][1].shape[0] == hi - lo
    # stand-in for the per-rank result map: mean over channels of the shard's keyframe
    res = mine["keyframe"].mean(1, keepdim=True)
    full = all_gather_batch(res)
    ok = torch.equal(full, data["keyframe"].mean(1, keepdim=True))
    if batch % world == 0:   # the no-size-exchange fast path gives the same tensor
        ok = ok and torch.equal(all_gather_batch(res, equal_shards=True), full)
    q.put((rank, bool(ok), tuple(full.shape)))
    dist.destroy_process_group()


def _run(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(ok for _, ok, _ in res), res
    assert all(shape[0] == batch for _, _, shape in res)


def test_even_shards_gloo():
    _run(4)


def test_ragged_shards_gloo():
    _run(5)


def test_shard_bounds_cover():
    from monorec_b200.dist import shard_bounds
    for batch in (1, 7, 8, 128):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
