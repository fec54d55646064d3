"""CPU (gloo, world_size 2): the data-parallel plumbing of diffusionmodel_b200/parallel.py.

The reference has no distributed code (new_scripy.py:676 hard-codes cuda:0); data parallelism is the
north-star's addition, so what is pinned here is its contract: the all-reduced
flat gradient equals the mean of the per-rank gradients, every rank ends with identical parameters
after rank 0's broadcast, sampling shards are class-aligned and cover the request, and the timing
reduction is a max over ranks.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusionmodel_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, local, w = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and parallel.world_size() == world
    # parameters: rank 0 wins
    torch.manual_seed(100 + rank)
    flat_param = torch.randn(1000)
    buf = torch.randn(7)
    parallel.broadcast_parameters(flat_param, [buf])
    # gradients: mean over ranks, whole buffer and bucketed
    torch.manual_seed(200 + rank)
    g = torch.randn(1000)
    mine = g.clone()
    a = parallel.allreduce_mean_(g.clone())
    b = parallel.allreduce_mean_(g.clone(), bucket_elems=128)
    t = parallel.max_over_ranks(10.0 + rank, torch.device("cpu"))
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    out[rank] = dict(param=flat_param, buf=buf, a=a, b=b, t=t, ref=torch.stack(gathered).mean(0),
                     shard=parallel.shard_samples(35, 5, rank, world))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_broadcast_and_shards():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = {k: v for k, v in out.items()}
    assert torch.equal(res[0]["param"], res[1]["param"]) and torch.equal(res[0]["buf"], res[1]["buf"])
    for r in range(world):
        assert torch.allclose(res[r]["a"], res[r]["ref"], atol=1e-6)
        assert torch.allclose(res[r]["b"], res[r]["ref"], atol=1e-6)
        assert res[r]["t"] == 11.0                              # max over ranks
    assert torch.equal(res[0]["a"], res[1]["a"])                # bit-identical on every rank
    shards = [res[r]["shard"] for r in range(world)]
    assert sum(shards) == 35 and all(s % 5 == 0 for s in shards)


def test_shard_samples_is_class_aligned_and_complete():
    for n_classes in (5, 10):
        for groups in (1, 3, 8, 16):
            n = groups * n_classes
            for world in (1, 2, 4, 8):
                parts = [parallel.shard_samples(n, n_classes, r, world) for r in range(world)]
                assert sum(parts) == n and all(p % n_classes == 0 for p in parts)
                assert max(parts) - min(parts) <= n_classes


def test_single_process_is_a_noop():
    g = torch.arange(8.0)
    assert parallel.world_size() == 1
    assert torch.equal(parallel.allreduce_mean_(g.clone()), g)
    assert parallel.max_over_ranks(3.5, torch.device("cpu")) == 3.5


class _FakeOpt:
    """What OverlappedGradReduce needs from FusedAdamW: parameter list, flat offsets, the flat gradient, flush()."""

    def __init__(self, sizes):
        self._params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
        self._offsets, off = [], 0
        for n in sizes:
            self._offsets.append(off)
            off += (n + 7) // 8 * 8
        self._n = off
        self.flat_grad = torch.zeros(off)
        self.flushes = 0

    def flush(self):
        self.flushes += 1


def _overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_from_env(backend="gloo")
    opt = _FakeOpt([40, 13, 64, 5, 100, 24])            # "trunk" = params 0-2 (the middle region = param 2), "decoder side" = 3-5
    p = opt._params
    red = parallel.OverlappedGradReduce(opt, [(p[3], None), (p[2], p[3])])
    res = []
    for step in range(2):                               # twice: the bookkeeping resets between optimizer steps
        torch.manual_seed(10 * step + rank)
        g = torch.randn(opt._n)
        gathered = [torch.zeros(opt._n) for _ in range(world)]
        dist.all_gather(gathered, g)
        want = torch.stack(gathered).mean(0)
        opt.flat_grad.copy_(g)
        red.reduce_ready(0)                             # decoder-side region: final first
        red.reduce_ready(1)                             # then the middle region
        red.finish()                                    # the rest + join
        res.append((opt.flat_grad.clone(), want))
    # a step without any staged region: finish() alone reduces everything
    torch.manual_seed(99 + rank)
    g = torch.randn(opt._n)
    gathered = [torch.zeros(opt._n) for _ in range(world)]
    dist.all_gather(gathered, g)
    opt.flat_grad.copy_(g)
    red.finish()
    res.append((opt.flat_grad.clone(), torch.stack(gathered).mean(0)))
    out[rank] = dict(res=res, spans=red.spans, flushes=opt.flushes)
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_grad_reduce_regions_cover_the_buffer_once():
    """parallel.OverlappedGradReduce on two gloo ranks: regions all-reduced as they become final plus the remainder in
    finish() give exactly the mean of the per-rank gradients -- every element reduced once, none twice."""
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_overlap_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = {k: v for k, v in out.items()}
    assert res[0]["spans"] == [(120, 256), (56, 120)]            # [offset(p3), n) and [offset(p2), offset(p3)): 8-aligned slots
    for r in range(world):
        for got, want in res[r]["res"]:
            assert torch.allclose(got, want, atol=1e-6)
        assert res[r]["flushes"] == 2 * 3 + 1
    assert torch.equal(res[0]["res"][0][0], res[1]["res"][0][0])
