"""CPU suite: the N>1 host logic (batch sharding, max-over-ranks timing) on gloo, world_size 2."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hydranet_b200.sharding import max_over_ranks, shard_bounds


def test_shard_bounds_partition():
    for gb in (0, 1, 7, 32, 256, 257):
        for w in (1, 2, 3, 8):
            cuts = [shard_bounds(gb, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_bounds(9, world, rank)
    x = torch.arange(9.0)[b:e] * 2  # stand-in for a per-image result: images are independent
    parts = [None] * world
    dist.all_gather_object(parts, x.tolist())
    t = max_over_ranks(1.0 + rank, dist)
    q.put((rank, sum(parts, []), t))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, t in res:
        assert gathered == (torch.arange(9.0) * 2).tolist()  # shards reassemble to the single-process result
        assert t == 2.0  # max over ranks
