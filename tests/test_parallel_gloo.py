"""world_size-2 gloo test (CPU) of the episode-sharding plumbing used for N > 1 GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from interactron_b200 import parallel
    ids = list(range(100, 111))                       # 11 episodes: ragged over 2 ranks
    mine = parallel.shard_episodes(ids)
    local = [(e, {"episode": e, "score": float(e) * 0.5}) for e in mine]
    allres = parallel.gather_detections(local, len(ids))
    tmax = parallel.max_over_ranks([1.0 + rank, 5.0 - rank])
    # meta-training step exchanges: SUM all-reduce of the flat meta-gradient, ordered path replay
    flat = torch.full((1, 7), float(rank + 1))
    parallel.allreduce_meta_grads(flat)
    order = parallel.exchange_in_order([("ep%d" % e, [rank, e % 4], 0.5 * e) for e in mine])
    # the two-bucket exchange forward() uses: [theta | psi | phi] with phi reduced first
    G = torch.arange(10, dtype=torch.float32).view(1, 10) * (rank + 1)
    red = parallel.BucketedAllReduce()
    red.launch_async(G[:, 6:])
    red.finish(G[:, :6])
    dist.barrier()
    q.put((rank, mine, [r["episode"] for r in allres], tmax, flat.tolist(), order, G.tolist(), red.ms()))
    dist.destroy_process_group()


def test_episode_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = list(range(100, 111))
    assert res[0][1] == ids[0::2] and res[1][1] == ids[1::2]           # disjoint cover, round robin
    for _, _, gathered, tmax, flat, order, G, ar_ms in res:
        assert G == [[3.0 * i for i in range(10)]]                      # both buckets summed over the ranks
        assert tuple(ar_ms) == (0.0, 0.0)                               # CPU tensors: no CUDA events
        assert gathered == ids                                          # order restored on every rank
        assert tmax == [2.0, 5.0]                                       # max over ranks
        assert flat == [[3.0] * 7]                                      # 1 + 2 summed on both ranks
        assert [it[0] for _, _, it in order] == ["ep%d" % e for e in ids]   # global episode order
        assert [(r, j) for r, j, _ in order] == [(i % 2, i // 2) for i in range(len(ids))]


def test_single_process_fallbacks():
    from interactron_b200 import parallel
    assert parallel.shard_episodes([1, 2, 3]) == [1, 2, 3]
    assert parallel.gather_detections([(1, "a"), (2, "b")], 2) == ["a", "b"]
    assert parallel.max_over_ranks([3.0]) == [3.0]
    assert parallel.exchange_in_order(["a", "b"]) == [(0, 0, "a"), (0, 1, "b")]
    t = torch.ones(1, 3)
    assert parallel.allreduce_meta_grads(t) is t
    red = parallel.BucketedAllReduce()
    red.launch_async(t[:, 2:])
    red.finish(t[:, :2])
    assert t.tolist() == [[1.0, 1.0, 1.0]] and red.ms() == (0.0, 0.0)
