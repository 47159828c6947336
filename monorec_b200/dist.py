"""Batch sharding for the one-process-per-GPU layout (replaces the reference's single-process
torch.nn.DataParallel scatter/gather: base/base_trainer.py:26-29, evaluater/evaluater.py:29-30).

Every keyframe is independent end to end, so the path shards on dim 0 with no data-path collective;
the only exchange is the optional all-gather of the per-rank result maps.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous, balanced [lo, hi) of a batch of `batch` keyframes for `rank` of `world`."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_data_dict(data, rank, world):
    """Slices every tensor (and every tensor in a list) of a MonoRec data_dict on dim 0."""
    batch = data["keyframe"].shape[0]
    lo, hi = shard_bounds(batch, rank, world)
    out = {}
    for k, v in data.items():
        if isinstance(v, (list, tuple)):
            out[k] = [t[lo:hi] if torch.is_tensor(t) and t.dim() > 0 and t.shape[0] == batch else t for t in v]
        elif torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == batch:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def all_gather_batch(t, group=None, equal_shards=False):
    """All-gather of per-rank (b_r, ...) maps into (sum b_r, ...) on every rank (one ncclAllGather when the shards
    are equal, a padded gather otherwise).  `equal_shards=True` asserts the caller knows every rank holds the same b_r
    (a batch divisible by the world size) and skips the size exchange and its host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    if equal_shards:
        out = t.new_empty((world * t.shape[0],) + tuple(t.shape[1:]))
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    sizes = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [int(s.item()) for s in all_sizes]
    if len(set(all_sizes)) == 1:
        out = t.new_empty((world * t.shape[0],) + tuple(t.shape[1:]))
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    mx = max(all_sizes)
    pad = t.new_zeros((mx,) + tuple(t.shape[1:]))
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, all_sizes)], 0)
