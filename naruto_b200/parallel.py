"""Host-side pieces of the ray-sharded data-parallel mapping iteration.

The path shards over rays: every op up to the per-ray outputs is row-wise in the ray batch, parameters are replicated.
Two exchanges per iteration, both plain all-reduces issued on the compute stream:
  1. the loss statistics (NRT_N_STATS_SUM fp64 sums/counts) before the backward pass -- each loss is a ratio of
     GLOBAL sums (tp/model/utils.py:103-107,121-122; src/slam/coslam/model/scene_rep.py:253-259,284);
  2. the flat fp32 gradient bucket [grid | w1 | w2 | w3 | w4 | uncert] before the (replicated) Adam step.
These helpers are backend-agnostic (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

from ._lib import N_STATS_SUM


def shard_range(n_rays: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of a global ray batch for `rank`; the slices tile [0, n_rays)."""
    base, rem = divmod(n_rays, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_stats(stats: torch.Tensor, group=None):
    """Sum the additive loss statistics across ranks in place (entries >= N_STATS_SUM are rank-local scratch)."""
    if group is not None and dist.get_world_size(group) > 1:
        dist.all_reduce(stats[:N_STATS_SUM], op=dist.ReduceOp.SUM, group=group)
    return stats


def reduce_grads(flat_grad: torch.Tensor, group=None):
    """Sum the flat gradient bucket across ranks in place.  No averaging: the per-rank gradients are already partial
    sums of the gradient of the GLOBAL-mean losses (they were formed with the globally reduced statistics)."""
    if group is not None and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad
