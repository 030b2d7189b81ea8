"""GPU: the fused mapping iteration (naruto_b200/mapper.py: render fwd, losses, backward, smoothness, Adam incl. the
every-5th uncertainty-grid step) against the oracle's mapping_iteration (torch autograd + torch.optim.Adam) driven
with the same uniform draws, eager and as a replayed CUDA graph; plus kernel-level checks of smoothness and Adam."""
import numpy as np
import pytest
import torch

from oracle import naruto_oracle as no
from oracle.make_golden import synth_rays

pytestmark = pytest.mark.gpu


def _setup(dev, B, n_samples_d, seed, use_graph):
    from naruto_b200.configs import replica_office0
    from naruto_b200.field import FieldPlan, FieldTensors
    from naruto_b200.mapper import MappingStep
    sp = no.office0_spec(n_samples_d=n_samples_d)
    P = no.init_params(sp, seed=seed, grid_range=0.05, uncert_jitter=0.5)
    cfg = replica_office0(n_samples_d=n_samples_d)
    plan = FieldPlan(cfg, sp.bound)
    ms = MappingStep(plan, cfg, B, dev, init=FieldTensors(P.grid, P.w1, P.w2, P.w3, P.w4, P.uncert_grid), use_graph=use_graph)
    ms.external_random = True
    return sp, P, ms


@pytest.mark.parametrize('use_graph', [False, True])
def test_mapping_iterations_match_oracle(use_graph):
    dev = torch.device('cuda:0')
    B, iters = 96, 6                                  # iteration 5 takes the uncertainty-grid Adam step
    sp, P, ms = _setup(dev, B, 32, 41, use_graph)
    Po = P.clone(requires_grad=True)
    opt = no.MappingOptimisers(Po)
    g = torch.Generator().manual_seed(77)
    for it in range(iters):
        o, d, rgb, td = synth_rays(sp, B, seed=500 + it)
        u = torch.rand(B, sp.n_samples, generator=g)
        r6 = torch.rand(6, generator=g)
        loss_o, ret_o = no.mapping_iteration(Po, opt, o, d, rgb, td, sp, it, u=u, smooth_draws=(r6[:3], r6[3:].view(1, 1, 1, 3)))
        ms.u.copy_(u)
        ms.rand6.copy_(r6)
        losses = ms.step(o.to(dev), d.to(dev), rgb.to(dev), td.to(dev))
        torch.cuda.synchronize()
        for k, name in enumerate(('rgb_loss', 'depth_loss', 'sdf_loss', 'fs_loss', 'uncert_loss')):
            a, b = losses[k].item(), ret_o[name].item()
            assert abs(a - b) <= 5e-4 * abs(b) + 1e-7, (it, name, a, b)
        assert abs(ms.total_loss() - loss_o.item()) <= 5e-4 * abs(loss_o.item()), (it, ms.total_loss(), loss_o.item())
    # parameters after 6 Adam steps.  Adam with eps=1e-15 turns a gradient into ~lr*sign(g), so entries whose
    # gradient is at rounding-noise level may step the other way: compare distributions, not every entry.
    for name, a, b, tol, frac in (('w1', ms.P.w1, Po.w1, 2e-3, 0.99), ('w2', ms.P.w2, Po.w2, 2e-3, 0.99),
                                  ('w3', ms.P.w3, Po.w3, 2e-3, 0.99), ('w4', ms.P.w4, Po.w4, 2e-3, 0.99),
                                  ('uncert', ms.P.uncert, Po.uncert_grid, 2e-2, 0.99),
                                  ('grid', ms.P.grid, Po.grid, 2e-3, 0.995)):
        d_ = (a.detach().cpu() - b.detach()).abs().reshape(-1)
        ok = (d_ <= tol).float().mean().item()
        assert ok >= frac, f'{name}: {ok * 100:.3f}% of entries within {tol} (max {d_.max():.3e})'
    # the uncertainty grid must have moved (lr=1 Adam step at iteration 5) and only then
    assert (ms.P.uncert.cpu() - P.uncert_grid).abs().max() > 0.1
    assert ms.launches_per_iter[False] == 8 and ms.launches_per_iter[True] == 9      # (external draws: one counter launch more on uncertainty steps)


@pytest.mark.parametrize('use_graph', [False, True])
def test_device_side_random_draws_change_every_iteration(use_graph):
    """Production mode (no caller-provided draws): the stratified jitter and the smoothness-lattice offsets come from Philox
    keyed by (seed, device step counter), so a REPLAYED graph still draws new values each iteration; depths stay sorted and
    inside the strata; two ranks (seeds) draw different jitter."""
    dev = torch.device('cuda:0')
    sp, P, ms = _setup(dev, 64, 32, 5, use_graph)
    ms.external_random = False
    o, d, rgb, td = synth_rays(sp, 64, seed=77)
    zs, r6 = [], []
    for it in range(4):
        losses = ms.step(o.to(dev), d.to(dev), rgb.to(dev), td.to(dev))
        torch.cuda.synchronize()
        assert torch.isfinite(losses[:5]).all()
        z = ms.out.z_vals.clone()
        assert (z[:, 1:] >= z[:, :-1]).all()
        zs.append(z)
        r6.append(ms.rand6.clone())
        assert (ms.rand6 >= 0).all() and (ms.rand6 < 1).all()
    for a in range(4):
        for b in range(a + 1, 4):
            assert (zs[a] != zs[b]).float().mean() > 0.9, 'jitter must change between iterations'
            assert not torch.equal(r6[a], r6[b])
    # jitter is uniform inside each stratum: the mean offset from the unperturbed depths is ~0 over many samples
    ms.seed += 7919                                   # what another rank would use
    ms._graphs.clear()
    ms.step(o.to(dev), d.to(dev), rgb.to(dev), td.to(dev))
    torch.cuda.synchronize()
    assert (ms.out.z_vals != zs[-1]).float().mean() > 0.9


def test_smoothness_vs_oracle():
    dev = torch.device('cuda:0')
    sp, P, ms = _setup(dev, 8, 32, 43, False)
    r6 = torch.rand(6, generator=torch.Generator().manual_seed(5))
    Pg = P.clone(requires_grad=True)
    sm = no.smoothness(Pg, sp, r6[:3], r6[3:].view(1, 1, 1, 3))
    (2.5 * sm).backward()
    ms.rand6.copy_(r6)
    ms.grad.zero_()
    ms.plan.smooth_fwd_bwd(ms.P.grid, ms.rand6, sp.smooth_pts, sp.smooth_vox, sp.smooth_margin, 2.5, ms.smooth_loss, ms.G.grid,
                           ms.ws_smooth)
    torch.cuda.synchronize()
    assert abs(ms.smooth_loss.item() - sm.item()) <= 1e-4 * abs(sm.item())
    ga, gb = ms.G.grid.cpu(), Pg.grid.grad
    assert (ga - gb).abs().max() <= 1e-4 * gb.abs().max()


def test_smoothness_slabs_add_up_to_the_whole_term():
    """Data-parallel ranks split the ray-independent smoothness term into slabs of the lattice (part / n_parts): the slabs'
    losses and gradients must add up to the single-GPU term."""
    dev = torch.device('cuda:0')
    sp, P, ms = _setup(dev, 8, 32, 43, False)
    ms.rand6.copy_(torch.rand(6, generator=torch.Generator().manual_seed(6)))
    args = (ms.P.grid, ms.rand6, sp.smooth_pts, sp.smooth_vox, sp.smooth_margin, 1.5)
    ms.grad.zero_()
    ms.plan.smooth_fwd_bwd(*args, ms.smooth_loss, ms.G.grid, ms.ws_smooth)
    torch.cuda.synchronize()
    whole_loss, whole_grad = ms.smooth_loss.item(), ms.G.grid.clone()
    assert whole_loss > 0 and whole_grad.abs().max() > 0
    for n_parts in (2, 3, 8):
        ms.grad.zero_()
        tot = 0.0
        for part in range(n_parts):
            ms.plan.smooth_fwd_bwd(*args, ms.smooth_loss, ms.G.grid, ms.ws_smooth, part=part, n_parts=n_parts)
            torch.cuda.synchronize()
            tot += ms.smooth_loss.item()
        assert abs(tot - whole_loss) <= 1e-5 * whole_loss, (n_parts, tot, whole_loss)
        assert (ms.G.grid - whole_grad).abs().max() <= 1e-5 * whole_grad.abs().max(), n_parts


def test_adam_kernel_vs_torch():
    dev = torch.device('cuda:0')
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.field import FieldPlan
    plan = FieldPlan(replica_office0(), OFFICE0_BOUND)
    g = torch.Generator().manual_seed(3)
    n = 100003
    for kw in (dict(lr=0.01, betas=(0.9, 0.99), eps=1e-15, weight_decay=0.0), dict(lr=0.01, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-6),
               dict(lr=1.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)):
        p0 = torch.randn(n, generator=g)
        pt = p0.clone().requires_grad_(True)
        opt = torch.optim.Adam([pt], **kw)
        pk, m, v = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        for step in range(1, 5):
            grad = torch.randn(n, generator=g) * (10.0 ** torch.randint(-8, 1, (n,), generator=g).float())
            grad[::7] = 0.0
            pt.grad = grad.clone()
            opt.step()
            gk = grad.to(dev)
            if step % 2:
                plan.adam_step(pk, gk, m, v, step, kw['lr'], kw['betas'][0], kw['betas'][1], kw['eps'], kw['weight_decay'], zero_grad=True)
            else:   # device-resident step counter (graph replay path)
                step_dev.fill_(step)
                plan.adam_step(pk, gk, m, v, 0, kw['lr'], kw['betas'][0], kw['betas'][1], kw['eps'], kw['weight_decay'], zero_grad=True,
                               step_dev=step_dev)
            assert (gk == 0).all()
            err = (pk.cpu() - pt.detach()).abs().max().item()
            assert err <= 2e-6 * max(1.0, kw['lr'] * 100), (kw, step, err)


def test_grouped_adam_equals_per_group_adam_and_skips_disabled_groups():
    """nrt_adam_step_groups (the iteration's one optimiser launch) against nrt_adam_step group by group: same bits, a ragged
    last group (scalar tail), zero_grad, and a disabled group left untouched with its gradient still accumulating."""
    dev = torch.device('cuda:0')
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.field import FieldPlan
    plan = FieldPlan(replica_office0(), OFFICE0_BOUND)
    g = torch.Generator().manual_seed(11)
    sizes = [40000, 5184, 9999 + 2]                     # begins are multiples of 4, the last end is not
    total = sum(sizes)
    hp = [(0.01, 0.9, 0.99, 1e-15, 0.0), (0.01, 0.9, 0.99, 1e-8, 1e-6), (1.0, 0.9, 0.999, 1e-8, 0.0)]
    pa = torch.randn(total, generator=g).to(dev)
    pb = pa.clone()
    ma, va, mb, vb = (torch.zeros(total, device=dev) for _ in range(4))
    steps = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(2)]
    ga_acc = torch.zeros(total, device=dev)
    gb_acc = torch.zeros(total, device=dev)
    for it in range(1, 7):
        grad = (torch.randn(total, generator=g) * (10.0 ** torch.randint(-6, 1, (total,), generator=g).float())).to(dev)
        ga_acc += grad
        gb_acc += grad
        with_unc = it % 3 == 0
        plan.iteration_begin(steps[0], steps[1] if with_unc else None)
        assert int(steps[0].item()) == it and int(steps[1].item()) == it // 3
        off, groups = 0, []
        for k, n in enumerate(sizes):
            lr, b1, b2, eps, wd = hp[k]
            en = k < 2 or with_unc
            groups.append((off, off + n, lr, b1, b2, eps, wd, steps[0 if k < 2 else 1], en))
            if en:
                s = slice(off, off + n)
                plan.adam_step(pb[s], gb_acc[s], mb[s], vb[s], 0, lr, b1, b2, eps, wd, zero_grad=True, step_dev=steps[0 if k < 2 else 1])
            off += n
        plan.adam_step_groups(pa, ga_acc, ma, va, groups, zero_grad=True)
        torch.cuda.synchronize()
        assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb), it
        assert torch.equal(ga_acc, gb_acc)
        if not with_unc:
            assert ga_acc[sizes[0] + sizes[1]:].abs().max() > 0 and (ga_acc[:sizes[0] + sizes[1]] == 0).all()
        else:
            assert (ga_acc == 0).all()


def test_step_host_is_the_same_iteration_fed_from_pinned_memory():
    """step_host (H2D memcpy node + iteration + D2H memcpy node in one graph) against load_packed + step + losses.cpu() on a
    twin: the first iteration's losses are bit-identical (same parameters, rays and jitter), later ones agree to the tolerance
    of the unordered table reductions."""
    dev = torch.device('cuda:0')
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.field import FieldPlan, FieldTensors
    from naruto_b200.mapper import MappingStep
    from naruto_b200.synthetic import SyntheticFrame
    cfg = replica_office0()
    plan = FieldPlan(cfg, OFFICE0_BOUND)
    g = torch.Generator().manual_seed(2)
    lin = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) / (i ** 0.5)
    init = FieldTensors((torch.rand(plan.n_grid_floats, generator=g) * 2 - 1) * 1e-2, lin(32, 80), lin(16, 32), lin(32, 63), lin(3, 32),
                        torch.full(plan.uncert_dims, 3.0))
    B = 300
    a = MappingStep(plan, cfg, B, dev, init=init)
    b = MappingStep(plan, cfg, B, dev, init=init)
    frame = SyntheticFrame(OFFICE0_BOUND, seed=4)
    for it in range(6):
        host = frame.sample_packed(B, pin=True)
        a.load_packed(host)
        la = a.step().cpu()
        lb = b.step_host(host)
        torch.cuda.synchronize()
        if it == 0:
            assert torch.equal(la, lb), (la, lb)
        assert torch.allclose(la[:5], lb[:5], rtol=2e-3, atol=1e-7), (it, la, lb)
    assert int(a.state.map_step.item()) == int(b.state.map_step.item()) == 6 and int(b.state.unc_step.item()) == 1
