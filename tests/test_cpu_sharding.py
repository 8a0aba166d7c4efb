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


# ------------------------------------------------------------------ training: bucketed gradient all-reduce
def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hydranet_b200.parallel import GradAllReduce
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    unused = torch.nn.Parameter(torch.ones(3))  # like neck.bifpn.0.p5_to_p6.*: registered, never in the graph
    params = list(net.parameters()) + [unused]
    red = GradAllReduce(params, bucket_mb=0.0001)  # ~100 B buckets: several buckets, launched from the hooks in reverse order
    assert len(red.buckets) >= 3 and red.buckets[0][0] is unused, [len(b) for b in red.buckets]
    outs = []
    for step in range(2):
        for p in params:
            p.grad = None
        x = torch.full((5, 16), float(rank + 1 + step))
        net(x).sum().backward()
        launched_by_hooks = sum(red.launched)
        red.finish()
        outs.append([None if p.grad is None else p.grad.clone() for p in params])
    q.put((rank, launched_by_hooks, outs))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_matches_mean_of_local_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=180) for _ in ps), key=lambda r: r[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth: mean over the two ranks' local gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    for step in range(2):
        want = None
        for rank in range(2):
            net.zero_grad()
            net(torch.full((5, 16), float(rank + 1 + step))).sum().backward()
            g = [p.grad.clone() for p in net.parameters()]
            want = g if want is None else [a + b for a, b in zip(want, g)]
        want = [w / 2 for w in want]
        for rank, launched, outs in res:
            got = outs[step]
            assert got[-1] is None  # the unused parameter stays without gradient on every rank
            for a, b in zip(got[:-1], want):
                assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
            assert launched >= 1  # buckets were launched from the autograd hooks, before finish()
