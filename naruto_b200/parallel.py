"""Host-side pieces of the ray-sharded data-parallel mapping iteration.

The path shards over rays: every op up to the per-ray outputs is row-wise in the ray batch, parameters are replicated.
Two exchanges per iteration, both plain all-reduces issued on the compute stream:
  1. the loss statistics (NRT_N_STATS_SUM fp64 sums/counts) before the backward pass -- each loss is a ratio of
     GLOBAL sums (tp/model/utils.py:103-107,121-122; src/slam/coslam/model/scene_rep.py:253-259,284);
  2. the flat fp32 gradient bucket [grid | w1 | w2 | w3 | w4 | uncert] before the (replicated) Adam step.
These helpers are backend-agnostic (NCCL on the GPUs, gloo in the CPU tests).
"""
import os

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


class PeerExchange:
    """NVLink peer-memory exchanges for the data-parallel iteration (csrc/peer.cu): ONE symmetric allocation per rank holding
    [theta | bucket | stats pads | flags], mapped by every peer (torch.distributed._symmetric_memory does the handle
    exchange over the process group).  `stats_exchange` replaces all-reduce(stats) + loss_finalize, `adam_step` replaces
    all-reduce(bucket) + the Adam launches.  Raises at construction if symmetric memory is not available for the group
    (callers fall back to the NCCL path of reduce_stats / reduce_grads)."""

    def __init__(self, total: int, device, group):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm

        from . import _lib as L
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError('PeerExchange: at most 8 ranks (one NVSwitch domain)')
        n_stats = 16
        pad4 = lambda n: (n + 3) // 4 * 4
        self.total = total
        off_theta = 0
        off_bucket = pad4(total)
        # gradients [0, total), then (behind the last float4 of the gradients) the smoothness-loss slot
        self.smooth_slot = pad4(total)
        off_stats = off_bucket + pad4(total) + 4               # floats; fp64 region must be 8-byte aligned (it is: multiples of 4)
        off_flags = off_stats + 2 * self.world * 2 * n_stats      # u64 [world][2 * n_stats]: flag-carrying words (csrc/peer.cu)
        n_floats = off_flags + 32
        name = group.group_name
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            symm.enable_symm_mem_for_group(name)      # a no-op on recent torch, required on older ones
        self.buf = symm.empty(n_floats, dtype=torch.float32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert len(ptrs) == self.world
        self.theta = self.buf[off_theta:off_theta + pad4(total)]        # padded: the kernels move float4s
        self.bucket = self.buf[off_bucket:off_bucket + pad4(total) + 4]
        self.stats_pad = self.buf[off_stats:off_flags]
        self.flags = self.buf[off_flags:off_flags + 32]
        t = L.NrtPeerTable()
        t.world, t.rank = self.world, self.rank
        for r, base in enumerate(ptrs):
            t.theta[r] = base + 4 * off_theta
            t.bucket[r] = base + 4 * off_bucket
            t.stats_pad[r] = base + 4 * off_stats
            t.flags[r] = base + 4 * off_flags
        # NVSwitch multicast mapping of the same allocation (NVLS), when the system provides one: in-switch gradient reduction.
        # Opt-in (NRT_DP_MULTICAST=1): measured on 8 B200s it shortens the optimiser launch's slice phase (25.4 -> 21.1 us) but the
        # iteration as a whole came out slower (477 vs 467 us, profiles/r02d_dp_stages.log); at 2 ranks it doubles the traffic.
        mc = int(getattr(self.handle, 'multicast_ptr', 0) or 0)
        self.multicast = bool(mc) and os.environ.get('NRT_DP_MULTICAST', '0') == '1'
        if self.multicast:
            t.bucket_mc, t.theta_mc = mc + 4 * off_bucket, mc + 4 * off_theta
        self.table = t
        self.xchg = torch.zeros(1, dtype=torch.int32, device=device)       # exchange counter (never restored by graph warm-ups)
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        self.lib = L.load()
        self._C, self._L = C, L
        dist.barrier(group)            # every rank has zero-filled its buffer before anybody raises a flag in it

    def stats_exchange(self, stats, losses):
        C, L = self._C, self._L
        L.check(self.lib.nrt_stats_exchange(C.byref(self.table), L.ptr(stats), L.ptr(self.xchg), L.ptr(losses),
                                            torch.cuda.current_stream().cuda_stream))

    def adam_step(self, exp_avg, exp_avg_sq, groups, smooth_total, keep_grad=None):
        """groups: list of (begin, end, lr, beta1, beta2, eps, weight_decay, step_dev tensor, enabled); keep_grad: per group, True =
        the gradients are not cleared on the peers (every rank clears its own bucket before it accumulates again)."""
        C, L = self._C, self._L
        arr = (L.NrtAdamGroup * len(groups))()
        for a, (b, e, lr, b1, b2, eps, wd, step_dev, en), keep in zip(arr, groups, keep_grad or [False] * len(groups)):
            a.begin, a.end, a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = b, e, lr, b1, b2, eps, wd
            a.step_dev, a.enabled, a.keep_grad = L.ptr(step_dev), int(en), int(keep)
        L.check(self.lib.nrt_adam_step_peers(C.byref(self.table), L.ptr(exp_avg), L.ptr(exp_avg_sq), arr, len(groups), self.smooth_slot,
                                             L.ptr(smooth_total) if smooth_total is not None else None, L.ptr(self.xchg),
                                             L.ptr(self.done), torch.cuda.current_stream().cuda_stream))
