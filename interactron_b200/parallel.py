"""Episode sharding across the GPUs of one box (one process per GPU, torch.distributed).

Episodes are independent (the reference loops over them serially, models/interactron.py:84), so
evaluation shards them with NO data-path collective: rank r adapts episodes r, r+W, r+2W, ...
and the python-side detections are gathered once at the end (outside any timed region).  The only
collectives there are the barrier / max-over-ranks used for timing.

The meta-training step has the path's one real exchange: the batch's meta-gradient is the SUM over
episodes (reference models/interactron.py:123,134 call .backward() once per episode), so with the
episodes of a batch sharded over ranks it is ONE all-reduce(SUM) over the flat fp32 buffer
[theta | psi | phi] (NCCL over NVLink; `allreduce_meta_grads`).  `interactron.path_storage` - the trie
of best action paths that labels the policy loss (models/interactron.py:105-118) - is per-process
state updated by every episode, so the (initial_image_path, actions, reward) triples are exchanged
and replayed in global episode order on every rank (`exchange_in_order`)."""
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


def allreduce_meta_grads(flat):
    """In-place SUM over ranks of the flat meta-gradient buffer (one collective per step)."""
    _, w = world()
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


class BucketedAllReduce:
    """The meta-gradient exchange of one step in two buckets (reference semantics: gradients are summed over
    the tasks of the batch, models/interactron.py:121-134; engine/interactron_trainer.py:106-111 then clips
    and steps on the sum).  `launch_async(first)`: all-reduce(SUM) of the bucket that is final early, on a
    side stream that waits for the producer stream - the caller keeps launching compute.  `finish(second)`:
    all-reduce of the rest on the caller's stream, then join the side stream.  CUDA events bracket both
    collectives: `ms()` -> (first bucket, second bucket + join = the exposed part).  No-op without a
    process group (events still recorded, ~0 ms)."""

    _side = {}

    def __init__(self):
        self.ev = None

    @classmethod
    def _stream(cls, device):
        key = (device.type, device.index)
        if key not in cls._side:
            cls._side[key] = torch.cuda.Stream(device=device)
        return cls._side[key]

    def launch_async(self, buf):
        _, w = world()
        if not buf.is_cuda:                               # gloo / CPU tests: plain blocking collective
            if w > 1:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            return
        main = torch.cuda.current_stream(buf.device)
        side = self._stream(buf.device)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            e[0].record(side)
            if w > 1:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            e[1].record(side)
        buf.record_stream(side)
        self.ev = e

    def finish(self, buf):
        _, w = world()
        if not buf.is_cuda:
            if w > 1:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            return
        main = torch.cuda.current_stream(buf.device)
        e = self.ev
        e[2].record(main)
        if w > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        main.wait_event(e[1])
        e[3].record(main)

    def ms(self):
        """(first bucket on the side stream, exposed: second bucket + join) in ms; synchronises."""
        if self.ev is None:
            return 0.0, 0.0
        self.ev[3].synchronize()
        return self.ev[0].elapsed_time(self.ev[1]), self.ev[2].elapsed_time(self.ev[3])


def exchange_in_order(local_items):
    """local_items: this rank's per-episode host objects, local order.  Global episode i of a batch
    lives on rank i % W at local position i // W (shard_episodes).  -> list of (rank, local_pos, item)
    for ALL episodes of the batch in global order, identical on every rank."""
    r, w = world()
    if w == 1:
        return [(0, j, it) for j, it in enumerate(local_items)]
    buckets = [None] * w
    dist.all_gather_object(buckets, list(local_items))
    out, j = [], 0
    while any(j < len(b) for b in buckets):
        for rank, b in enumerate(buckets):
            if j < len(b):
                out.append((rank, j, b[j]))
        j += 1
    return out
