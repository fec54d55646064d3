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
