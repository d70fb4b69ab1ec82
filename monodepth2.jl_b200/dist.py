"""Batch-sharded data parallelism for the loss path: one process per GPU, every
rank runs the same kernels on its contiguous chunk of the batch, nothing on the data path is
exchanged.  Only scalars (timings, the loss for logging) and -- in a training step -- the
parameter gradients cross ranks."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """contiguous, as-equal-as-possible chunk [lo, hi) of a batch for `rank`"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank: int, world: int):
    """slice every tensor (or list of tensors) of a batch along dim 0"""
    def cut(t):
        if isinstance(t, (list, tuple)):
            return type(t)(cut(u) for u in t)
        lo, hi = shard_bounds(t.shape[0], rank, world)
        return t[lo:hi]
    return cut(tensors)


def global_loss(local_loss: torch.Tensor, n_local: int) -> torch.Tensor:
    """mean over ALL images from per-rank means: sum_r n_r * loss_r / sum_r n_r (equals the mean
    of means for equal shards; src/training.jl:69 takes the mean over the whole batch)"""
    buf = torch.stack([local_loss.detach().double() * n_local, torch.tensor(float(n_local), dtype=torch.float64,
                                                                            device=local_loss.device)])
    if dist.is_initialized():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf[0] / buf[1]


def max_over_ranks(value_ms: float, device=None) -> float:
    """device-timed durations are reported as the max over ranks"""
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_mean_(grads, n_local=None):
    """in-place average of (parameter) gradients over ranks.  Each rank's gradient is that of the mean over ITS images,
    so with unequal shards (shard_bounds when N % world != 0) the full-batch mean of src/training.jl:69 is the
    n_local-weighted average: pass the rank's shard size as `n_local`; None means equal shards.
    (The training step uses the bucketed, overlapped GradientBuckets of train_step.py; this is the plain form.)"""
    if dist.is_initialized():
        w = dist.get_world_size()
        if n_local is None:
            scale_in, scale_out = 1.0, 1.0 / w
        else:
            tot = torch.tensor([float(n_local)], dtype=torch.float64, device=grads[0].device if grads else None)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            scale_in, scale_out = float(n_local) / float(tot.item()), 1.0
        for g in grads:
            if scale_in != 1.0:
                g.mul_(scale_in)
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            if scale_out != 1.0:
                g.mul_(scale_out)
    return grads
