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


NCCL_CTAS = 16          # SMs left to NCCL while a collective overlaps compute (DDPM.capture_train_step(trunk_sm_limit=148 - 16))


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            # the all-reduce that overlaps the trunk's backward shares the GPU with GEMM grids capped to 148 - 16 SMs
            os.environ.setdefault("NCCL_MAX_CTAS", str(NCCL_CTAS))
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


class OverlappedGradReduce:
    """The gradient all-reduce of one optimizer step in pieces, all but the last hidden behind the backward pass.

    The optimizer's flat gradient buffer follows ``parameters()`` order.  In the LAST micro-step of an accumulation window
    ``DDPM.capture_train_step(split_backward=True)`` replays the backward pass in stages (decoder side; down4 + CoordAttn;
    init_conv ... down3) and calls back after each: the regions of ``nn_model.grad_ready_regions()`` are final then -- at
    Cfg defaults 62 % of the bytes after stage 0 (up0 alone is 43 %) and 29 % after stage 1 -- so their all-reduces are
    issued there and run on NCCL's stream while the remaining stages execute; only the rest (9 %) stays exposed.
    Usage per optimizer step:

        last micro-step:  step(x, c, m, between=red.reduce_ready)
        then:             red.finish(); optimizer.step()
    """

    def __init__(self, optimizer, regions):
        self.opt = optimizer
        index = {id(p): i for i, p in enumerate(optimizer._params)}
        self.spans = []
        for first, end in regions:
            if id(first) not in index or (end is not None and id(end) not in index):
                raise ValueError("OverlappedGradReduce: a region boundary is not a parameter of the optimizer")
            lo = optimizer._offsets[index[id(first)]]
            hi = optimizer._offsets[index[id(end)]] if end is not None else optimizer._n
            self.spans.append((lo, hi))
        self._works, self._done = [], []

    @property
    def boundary(self):
        return self.spans[0][0]

    def _op(self):
        return dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM

    def reduce_ready(self, k=0):
        """Call when the gradients of region k are final (after backward stage k of the last micro-step)."""
        self.opt.flush()                      # packed / queued weight gradients (up0's deferred GEMM) -> flat buffer
        lo, hi = self.spans[k]
        self._done.append((lo, hi))
        if world_size() > 1:
            self._works.append(dist.all_reduce(self.opt.flat_grad[lo:hi], op=self._op(), async_op=True))

    reduce_tail = reduce_ready

    def finish(self):
        """After the last micro-step: reduce whatever no stage covered and join the overlapped collectives."""
        self.opt.flush()
        n = world_size()
        done, self._done = sorted(self._done), []
        if n == 1:
            return
        pos, rest = 0, []
        for lo, hi in done + [(self.opt._n, self.opt._n)]:
            if lo > pos:
                rest.append((pos, lo))
            pos = max(pos, hi)
        for lo, hi in rest:
            dist.all_reduce(self.opt.flat_grad[lo:hi], op=self._op())
        for w in self._works:
            w.wait()
        self._works = []
        if dist.get_backend() != "nccl":
            self.opt.flat_grad.mul_(1.0 / n)


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
