"""CPU, world_size 2, gloo: host-side logic of the ray-sharded data-parallel iteration (naruto_b200/parallel.py) and
the property it relies on -- per-shard gradients formed with GLOBALLY reduced loss statistics sum to the single-rank
gradient (checked with the oracle's arithmetic, since the kernels need a GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from naruto_b200 import _lib as L
        from naruto_b200.parallel import reduce_grads, reduce_stats, shard_range
        # 1. shard ranges tile the batch
        B = 77
        lo, hi = shard_range(B, rank, world)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([hi - lo]))
        assert sum(int(s) for s in sizes) == B and abs(int(sizes[0]) - int(sizes[1])) <= 1
        # 2. statistics: additive entries are summed, scratch entries stay local
        stats = torch.full((L.N_STATS + 4,), float(rank + 1), dtype=torch.float64)
        reduce_stats(stats, dist.group.WORLD)
        assert torch.all(stats[:L.N_STATS_SUM] == 3.0) and torch.all(stats[L.N_STATS_SUM:] == rank + 1)
        # 3. sharded loss with global normalisers == single-process loss; gradients sum
        torch.manual_seed(0)
        pred = torch.randn(B, 3)
        tgt = torch.randn(B, 3)
        w = torch.randn(3, 3, requires_grad=True)
        full = ((pred @ w - tgt) ** 2).mean()
        gfull, = torch.autograd.grad(full, w)
        w2 = w.detach().clone().requires_grad_(True)
        n_global = torch.tensor([float((hi - lo) * 3)], dtype=torch.float64)
        reduce_stats_like = n_global.clone()
        dist.all_reduce(reduce_stats_like)
        part = ((pred[lo:hi] @ w2 - tgt[lo:hi]) ** 2).sum() / reduce_stats_like.item()
        g, = torch.autograd.grad(part, w2)
        flat = g.reshape(-1).clone()
        reduce_grads(flat, dist.group.WORLD)
        assert torch.allclose(flat.view(3, 3), gfull, atol=1e-6)
        q.put((rank, 'ok'))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res


def test_shard_range_edges():
    from naruto_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_owned_ranges_tile_every_group():
    """The host-side ownership map of the sharded Adam moments (FusedState.owned_ranges) mirrors csrc/peer.cu: for every world
    size the ranks' ranges tile each parameter group exactly, in float4 units."""
    import types
    from naruto_b200.mapper import FusedState
    for n_grid, n_dec, n_unc in ((1628176, 5184, 96040), (38452256, 5184, 929016), (1628176, 5184, 993531)):
        st = types.SimpleNamespace(n_grid=n_grid, n_dec=n_dec, total=n_grid + n_dec + n_unc, peers=types.SimpleNamespace(world=1, rank=0))
        for world in (2, 3, 4, 8):
            cover = [0] * 3
            prev_hi = [0, n_grid, n_grid + n_dec]
            for rank in range(world):
                rs = FusedState.owned_ranges(st, rank=rank, world=world)
                for g, (lo, hi) in enumerate(rs):
                    assert lo == prev_hi[g] and lo % 4 == 0 and hi >= lo
                    prev_hi[g] = hi
            assert prev_hi == [n_grid, n_grid + n_dec, st.total]
