"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink; gloo in CPU tests).

The reference has no distributed code (single process, cuda:0: new_scripy.py:676).  Training shards by
data: full replica per rank, per-rank BatchNorm statistics (the reference has no SyncBN), one gradient
all-reduce per optimizer step on the optimizer's single flat gradient buffer, i.e. once per
ACCUM_STEPS micro-batches (new_scripy.py:795).  Sampling shards the independent trajectories with no
collective inside the 700-step loop.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def broadcast_parameters(flat_param: torch.Tensor, buffers=()):
    """Rank 0's parameters (one flat tensor) and buffers to every rank at start."""
    if world_size() == 1:
        return
    dist.broadcast(flat_param, 0)
    for b in buffers:
        dist.broadcast(b, 0)
    if flat_param.is_cuda:                       # parameters changed behind the weight-pack caches' back
        from . import ops
        ops.bump_weights_epoch()


def allreduce_mean_(flat_grad: torch.Tensor, bucket_elems: int = 0):
    """In-place mean of the flat gradient over ranks.  With NVSwitch the cost is latency- not
    link-bound, so the default is ONE collective over the whole buffer; ``bucket_elems`` > 0 splits it
    (async, all launched before the first wait) to overlap with whatever else is on the stream."""
    n = world_size()
    if n == 1:
        return flat_grad
    # NCCL averages inside the collective (no extra 2.8 GB scaling pass over the gradient); gloo only sums
    avg = dist.get_backend() == "nccl"
    op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
    if bucket_elems and bucket_elems < flat_grad.numel():
        works = [dist.all_reduce(flat_grad[o:o + bucket_elems], op=op, async_op=True)
                 for o in range(0, flat_grad.numel(), bucket_elems)]
        for w in works:
            w.wait()
    else:
        dist.all_reduce(flat_grad, op=op)
    if not avg:
        flat_grad.mul_(1.0 / n)
    return flat_grad


def shard_samples(n_sample: int, n_classes: int, rank: int, world: int):
    """Split ``n_sample`` trajectories (a multiple of n_classes, class-cycled like new_scripy.py:447-448)
    into per-rank counts that are themselves multiples of n_classes."""
    groups = n_sample // n_classes
    base, extra = divmod(groups, world)
    mine = base + (1 if rank < extra else 0)
    return mine * n_classes


def max_over_ranks(ms: float, device) -> float:
    if world_size() == 1:
        return ms
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
