"""GPU, needs >= 2 devices (skipped otherwise; run with `gpurun --gpus 2`, log under profiles/): the ray-sharded data-parallel
iteration on 2 ranks equals the single-GPU iteration on the concatenated batch -- losses, smoothness term, parameters after
three Adam steps (uncertainty-grid step on the third) -- for both exchange implementations: NVLink peer memory inside our own
kernels (csrc/peer.cu, the default) and NCCL all-reduces (NRT_DP_IMPL=nccl)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

B_GLOBAL, ITERS = 192, 3


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, impl, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), NRT_DP_IMPL=impl)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.field import FieldPlan, FieldTensors
    from naruto_b200.mapper import MappingStep
    from naruto_b200.parallel import shard_range
    from naruto_b200.synthetic import SyntheticFrame
    cfg = replica_office0(n_samples_d=32)
    plan = FieldPlan(cfg, OFFICE0_BOUND)
    g = torch.Generator().manual_seed(1)
    lin = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) / (i ** 0.5)
    init = FieldTensors((torch.rand(plan.n_grid_floats, generator=g) * 2 - 1) * 0.05, lin(32, 80), lin(16, 32), lin(32, 63), lin(3, 32),
                        3.0 + torch.rand(plan.uncert_dims, generator=g) - 0.5)
    lo, hi = shard_range(B_GLOBAL, rank, world)
    ms = MappingStep(plan, cfg, hi - lo, dev, init=init, process_group=dist.group.WORLD, use_graph=True)
    ms.external_random = True
    ref = None
    if rank == 0:
        ref = MappingStep(plan, cfg, B_GLOBAL, dev, init=init, use_graph=False)
        ref.external_random = True
    frame = SyntheticFrame(OFFICE0_BOUND, seed=3)
    gen = torch.Generator().manual_seed(7)
    log = {'impl': impl, 'peers': ms.peers is not None, 'losses': [], 'ref_losses': [], 'totals': [], 'ref_totals': []}
    for it in range(ITERS):
        o, d, rgb, td = frame.sample(B_GLOBAL)
        u = torch.rand(B_GLOBAL, plan.S, generator=gen)
        r6 = torch.rand(6, generator=gen)
        ms.u.copy_(u[lo:hi])
        ms.rand6.copy_(r6)
        ms.step(o[lo:hi].to(dev), d[lo:hi].to(dev), rgb[lo:hi].to(dev), td[lo:hi].to(dev), with_uncert_step=(it == ITERS - 1))
        torch.cuda.synchronize()
        log['losses'].append(ms.losses[:5].cpu())
        log['totals'].append(ms.total_loss())
        if ref is not None:
            ref.u.copy_(u)
            ref.rand6.copy_(r6)
            ref.step(o.to(dev), d.to(dev), rgb.to(dev), td.to(dev), with_uncert_step=(it == ITERS - 1))
            torch.cuda.synchronize()
            log['ref_losses'].append(ref.losses[:5].cpu())
            log['ref_totals'].append(ref.total_loss())
    ms.state.gather_moments()                 # peer memory shards the Adam moments: make them whole again (save_ckpt)
    log['theta'] = ms.theta.cpu()
    log['exp_avg'] = ms.exp_avg.cpu()
    if ref is not None:
        log['ref_theta'] = ref.theta.cpu()
        log['ref_exp_avg'] = ref.exp_avg.cpu()
        log['sizes'] = ms.state.sizes
    torch.save(log, os.path.join(out_dir, f'{impl}_{rank}.pt'))
    dist.barrier()
    ms.release_graphs()
    torch.cuda.synchronize()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
@pytest.mark.parametrize('impl', ['peer', 'nccl'])
def test_two_ranks_equal_one_rank_on_the_concatenated_batch(tmp_path, impl):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), impl, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / f'{impl}_0.pt'), torch.load(tmp_path / f'{impl}_1.pt')
    assert r0['peers'] == (impl == 'peer'), 'the requested exchange implementation must be the one that ran'
    # both ranks hold the same parameters, bit for bit (peer memory: each slice is reduced and stepped by ONE rank; NCCL: the
    # all-reduced bucket is identical everywhere and every rank takes the same Adam step)
    assert torch.equal(r0['theta'], r1['theta'])
    for it in range(ITERS):
        a, b = r0['losses'][it], r0['ref_losses'][it]
        assert torch.equal(r0['losses'][it], r1['losses'][it]), 'losses are ratios of global sums: identical on every rank'
        # iteration 0 starts from identical parameters: only summation order differs; later ones carry the Adam(eps=1e-15)
        # sign-like steps of entries whose gradient is at rounding-noise level (as in test_mapper.py)
        tol = 1e-5 if it == 0 else 5e-4
        assert ((a - b).abs() <= tol * b.abs() + 1e-9).all(), (it, a, b)
        assert abs(r0['totals'][it] - r0['ref_totals'][it]) <= tol * abs(r0['ref_totals'][it]), (it, r0['totals'][it], r0['ref_totals'][it])
    off = 0
    for name, n, tol, frac in zip(('grid', 'w1', 'w2', 'w3', 'w4', 'uncert'), r0['sizes'], (2e-3,) * 5 + (2e-2,), (0.995, 0.99, 0.99, 0.99, 0.99, 0.98)):   # uncert: one lr = 1 Adam step, sign-like for tiny gradients
        d = (r0['theta'][off:off + n] - r0['ref_theta'][off:off + n]).abs()
        ok = (d <= tol).float().mean().item()
        print(f'{impl} {name}: {ok:.4%} of entries within {tol} of the 1-GPU run (max {d.max().item():.2e})')
        assert ok >= frac, (name, ok)
        off += n
    assert (r0['theta'][-r0['sizes'][5]:] - 3.0).abs().max() > 0.1, 'the uncertainty grid took its Adam step'
    # the (gathered) first moments are complete and the same on both ranks, and they are the single-GPU run's
    assert torch.equal(r0['exp_avg'], r1['exp_avg'])
    dm = (r0['exp_avg'] - r0['ref_exp_avg']).abs()
    assert dm.max() <= 2e-3 * r0['ref_exp_avg'].abs().max(), (dm.max(), r0['ref_exp_avg'].abs().max())
