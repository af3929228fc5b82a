"""Episode sharding across the GPUs of one box (one process per GPU, torch.distributed).

Episodes are independent (the reference loops over them serially, models/interactron.py:84), so
evaluation shards them with NO data-path collective: rank r adapts episodes r, r+W, r+2W, ...
and the python-side detections are gathered once at the end (outside any timed region).  The only
collectives are the barrier / max-over-ranks used for timing.  (The meta-training step, not built
yet, is where the one real exchange — an NCCL all-reduce(SUM) of the meta-gradient — belongs.)
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_episodes(episode_ids, rank=None, world_size=None):
    """Round-robin shard: rank r gets episode_ids[r::W]."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(episode_ids)[rank::world_size]


def gather_detections(local, n_total):
    """local: list of (episode_id, payload) of this rank -> list of n_total payloads ordered by
    episode position (inverse of shard_episodes).  Uses all_gather_object (host objects)."""
    r, w = world()
    if w == 1:
        return [p for _, p in local]
    buckets = [None] * w
    dist.all_gather_object(buckets, local)
    out = [None] * n_total
    for rank, items in enumerate(buckets):
        for j, (_, payload) in enumerate(items):
            out[rank + j * w] = payload
    return out


def max_over_ranks(values, device="cpu"):
    """Element-wise max of a list of floats over ranks (timing reduction)."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    _, w = world()
    if w > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
