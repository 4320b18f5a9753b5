"""Multi-GPU plumbing: one process per GPU, batch-sharded, ONE collective.

Every op of the hot path is per-sample independent, so ranks process disjoint
batch shards with no data-path communication.  The only cross-sample quantities are the two
batch means behind get_prior_regularization_loss() / get_identity_metric(); each rank's kernel
leaves three floats [sum, sum, count] and a single all-reduce(sum) of that vector makes every rank
report the value the single-process reference computes on the un-sharded batch
(common/basecanonicalization.py:290-311, :390-430).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n samples for `rank`; the first n % world_size ranks get one extra."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int = None, world_size: int = None) -> torch.Tensor:
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(x.shape[0], rank, world_size)
    return x[lo:hi]


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the 3-float statistic over ranks (NCCL on CUDA tensors, gloo on CPU tensors in tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        stats = stats.clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def allreduce_stats_async(stats: torch.Tensor, group=None):
    """Start the sum of the 3-float statistic over ranks; -> (tensor, work).  `work.wait()` orders the current
    CUDA stream after the collective (NCCL) or blocks until it is done (gloo)."""
    out = stats.clone()
    work = dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return out, work


def mean_from_stats(stats: torch.Tensor, which: int, count_index: int, sync: bool = True, group=None) -> torch.Tensor:
    """stats[which] / stats[count_index], all-reduced first when `sync` and a process group is up."""
    s = allreduce_stats(stats, group) if sync else stats
    return s[which] / s[count_index]
