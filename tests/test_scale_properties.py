"""GPU: size-independent properties of the CUDA path at ragged and at full BASELINE sizes (the oracle is too slow there):
the fused ray kernel against the point kernel + stand-alone compositor, linearity / accumulation of the backward pass,
weight normalisation, empty inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _plan(n_samples_d, seed=3, grid_range=0.3):
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.field import FieldPlan, FieldTensors
    cfg = replica_office0(n_samples_d=n_samples_d)
    plan = FieldPlan(cfg, OFFICE0_BOUND)
    g = torch.Generator().manual_seed(seed)
    lin = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / (i ** 0.5)).cuda()
    P = FieldTensors(((torch.rand(plan.n_grid_floats, generator=g) * 2 - 1) * grid_range).cuda(), lin(32, 80), lin(16, 32),
                     lin(32, 63), lin(3, 32), (3.0 + torch.rand(plan.uncert_dims, generator=g) * 2 - 1).cuda())
    return cfg, plan, P


def _rays(B, seed):
    from naruto_b200.configs import OFFICE0_BOUND
    from naruto_b200.synthetic import SyntheticFrame
    o, d, rgb, td = SyntheticFrame(OFFICE0_BOUND, seed=seed).sample(B)
    return o.cuda(), d.cuda(), rgb.cuda(), td.cuda()


@pytest.mark.parametrize('n_samples_d,B', [(32, 1), (32, 5), (32, 301), (117, 3), (117, 130), (189, 7), (245, 2)])
def test_ray_kernel_equals_point_kernel_plus_compositor(n_samples_d, B):
    """render_fwd (fused) == decode_fwd on the same points followed by composite_fwd: same tensor-core tiles, so bit-exact
    `raw`; the per-ray outputs go through the same warp compositor."""
    from naruto_b200.field import RenderBuffers
    cfg, plan, P = _plan(n_samples_d)
    S = plan.S
    o, d, rgb, td = _rays(B, seed=B + S)
    u = torch.rand(B, S, generator=torch.Generator().manual_seed(1)).cuda()
    out = RenderBuffers(B, S, 'cuda', per_sample=True, weights=True, feat=True)
    plan.render_fwd(P, o, d, td, out, u=u)
    z = out.z_vals
    assert torch.all(z[:, 1:] >= z[:, :-1]), 'z_vals must be sorted'
    b = plan.bound.cuda()
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    x = ((pts - b[:, 0]) / (b[:, 1] - b[:, 0])).reshape(-1, 3).contiguous()
    raw, _, _ = plan.decode_fwd(P, x, with_color=True)
    assert torch.equal(raw.view(B, S, 5), out.raw), (raw.view(B, S, 5) - out.raw).abs().max()
    assert torch.equal(plan.encode_fwd(P.grid, x), out.feat_rows())
    comp = plan.composite_fwd(out.raw, z)
    for k in ('rgb', 'depth', 'depth_var', 'acc', 'disp', 'uncert', 'weights'):
        assert torch.equal(getattr(comp, k), getattr(out, k)), k
    torch.cuda.synchronize()
    w = out.weights
    assert torch.isfinite(out.rgb).all() and (w >= 0).all()
    assert torch.allclose(w.sum(1), out.acc, atol=1e-5) and (out.acc <= 1.0 + 1e-5).all()


@pytest.mark.parametrize('n_samples_d,B', [(32, 1), (32, 301), (117, 130), (117, 4096), (245, 9)])
def test_fused_loss_statistics_equal_stand_alone_kernel(n_samples_d, B):
    """nrt_render_fwd_stats (loss sums accumulated by the compositing warps) == nrt_render_fwd + nrt_loss_partial: identical
    outputs, integer counts exact, squared-error sums to fp32 summation-order noise; repeatable bit for bit."""
    from naruto_b200.field import RenderBuffers
    cfg, plan, P = _plan(n_samples_d)
    S = plan.S
    o, d, rgb, td = _rays(B, seed=B + S)
    u = torch.rand(B, S, generator=torch.Generator().manual_seed(1)).cuda()
    a = RenderBuffers(B, S, 'cuda', per_sample=True, feat=True)
    b = RenderBuffers(B, S, 'cuda', per_sample=True, feat=True)
    sa, sb, sc = plan.new_stats('cuda'), plan.new_stats('cuda'), plan.new_stats('cuda')
    plan.render_fwd(P, o, d, td, a, u=u)
    plan.loss_partial(a, rgb, td, sa)
    plan.render_fwd_stats(P, o, d, rgb, td, b, sb, u=u)
    plan.render_fwd_stats(P, o, d, rgb, td, b, sc, u=u)
    for k in ('rgb', 'depth', 'uncert', 'z_vals', 'raw', 'feat'):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    sa, sb, sc = sa[:12].cpu(), sb[:12].cpu(), sc[:12].cpu()
    assert torch.equal(sb, sc), 'fused statistics must be deterministic'
    for i in (0, 1, 2, 3, 4, 11):          # counts and the minimum: exact
        assert sa[i] == sb[i], (i, sa[i], sb[i])
    assert sb[0] == B and sb[4] == B * S
    for i in range(5, 11):
        assert abs(sa[i] - sb[i]) <= 2e-6 * abs(sa[i]) + 1e-12, (i, sa[i], sb[i])


def test_one_role_kernel_gives_identical_bits(tmp_path):
    """NRT_RENDER_IMPL=tc (the one-role thread-pair kernel kept for A/B runs) and the warp-specialised product kernel run the
    same arithmetic: outputs must be bit-identical.  The switch is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = '''
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from test_scale_properties import _plan, _rays
from naruto_b200.field import RenderBuffers
cfg, plan, P = _plan(117)
o, d, rgb, td = _rays(203, seed=9)
u = torch.rand(203, plan.S, generator=torch.Generator().manual_seed(1)).cuda()
out = RenderBuffers(203, plan.S, 'cuda', per_sample=True, weights=True, feat=True)
plan.render_fwd(P, o, d, td, out, u=u)
torch.save({k: getattr(out, k).cpu() for k in ('rgb', 'depth', 'uncert', 'z_vals', 'raw', 'weights', 'feat', 'masks')}, sys.argv[1])
''' % (root, os.path.join(root, 'tests'))
    res = {}
    for impl in ('ws', 'tc'):
        f = str(tmp_path / (impl + '.pt'))
        env = dict(os.environ, NRT_RENDER_IMPL=impl)
        subprocess.run([sys.executable, '-c', code, f], check=True, env=env, timeout=300)
        res[impl] = torch.load(f)
    for k in res['ws']:
        assert torch.equal(res['ws'][k], res['tc'][k]), k


def test_empty_inputs():
    from naruto_b200.field import RenderBuffers
    cfg, plan, P = _plan(32)
    e3 = torch.empty(0, 3, device='cuda')
    out = RenderBuffers(0, plan.S, 'cuda')
    plan.render_fwd(P, e3, e3, torch.empty(0, device='cuda'), out)
    raw, _, _ = plan.decode_fwd(P, e3)
    assert raw.shape == (0, 5)
    assert plan.encode_fwd(P.grid, e3).shape == (0, 32)
    torch.cuda.synchronize()


def test_backward_linearity_and_accumulation_at_bench_size():
    """4096 rays x 128 samples (the bench workload): d(2L) = 2 dL and two backward calls accumulate, for every parameter
    tensor; checked at the level the atomics' summation order allows."""
    from naruto_b200.field import FieldTensors, RenderBuffers
    cfg, plan, P = _plan(117, grid_range=0.05)
    B, S = 4096, plan.S
    o, d, rgb, td = _rays(B, seed=9)
    out = RenderBuffers(B, S, 'cuda', per_sample=True, feat=True)
    plan.render_fwd(P, o, d, td, out, perturb=0)
    stats = plan.new_stats('cuda')
    losses = torch.zeros(8, device='cuda')
    plan.loss_partial(out, rgb, td, stats)
    plan.loss_finalize(stats, losses)
    assert torch.isfinite(losses[:6]).all()
    lg = torch.tensor([5.0, 0.1, 1000.0, 10.0, 0.005], device='cuda')

    def grads(scale, calls):
        G = FieldTensors(*[torch.zeros_like(t) for t in P.as_list()])
        for _ in range(calls):
            plan.render_bwd(P, o, d, rgb, td, out, stats, lg * scale, G)
        torch.cuda.synchronize()
        return G

    g1, g2, gacc = grads(1.0, 1), grads(2.0, 1), grads(1.0, 2)
    for name, a, b, c in zip(('grid', 'w1', 'w2', 'w3', 'w4', 'uncert'), g1.as_list(), g2.as_list(), gacc.as_list()):
        s = a.abs().max().item()
        assert s > 0, name
        tol = 2e-3 if name.startswith('w') else 1e-4        # weight gradients: single-pass TF32 contraction over points
        assert (2 * a - b).abs().max().item() <= tol * 2 * s, name
        assert (2 * a - c).abs().max().item() <= tol * 2 * s, name


@pytest.mark.parametrize('n_rays,n_samples_d', [(1, 32), (7, 32), (301, 32), (96, 117), (4096, 117)])
def test_backward_q_kernel_matches_round1_kernel(tmp_path, n_rays, n_samples_d):
    """The 32-warp TMA-fed backward (backward_q.cu, the product path of nrt_render_bwd) against round 1's kernel
    (NRT_BWD_IMPL=tc) on the same saved forward, at ragged sizes (partial last tile, tiles spanning several rays) and at the
    bench shape.  Data gradients (hash table, uncertainty grid) come from the same 3xTF32 chain and may differ only by the
    summation order of the atomics; the MLP weight gradients are single-pass TF32 contractions over the points in both
    kernels, in different formulations (geo folded through W2 here).  The switch is read once per process: subprocesses."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = '''
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from test_scale_properties import _plan, _rays
from naruto_b200.field import FieldTensors, RenderBuffers
cfg, plan, P = _plan(%d, grid_range=0.05)
B = %d
o, d, rgb, td = _rays(B, seed=9)
out = RenderBuffers(B, plan.S, 'cuda', per_sample=True, feat=True)
stats = plan.new_stats('cuda')
losses = torch.zeros(8, device='cuda')
plan.render_fwd_stats(P, o, d, rgb, td, out, stats, u=torch.rand(B, plan.S, generator=torch.Generator().manual_seed(1)).cuda(), losses=losses)
lg = torch.tensor([5.0, 0.1, 1000.0, 10.0, 0.005], device='cuda')
G = FieldTensors(*[torch.zeros_like(t) for t in P.as_list()])
plan.render_bwd(P, o, d, rgb, td, out, stats, lg, G)
torch.cuda.synchronize()
torch.save({n: g.cpu() for n, g in zip(('grid', 'w1', 'w2', 'w3', 'w4', 'uncert'), G.as_list())}, sys.argv[1])
''' % (root, os.path.join(root, 'tests'), n_samples_d, n_rays)
    res = {}
    for impl in ('q', 'tc'):
        f = str(tmp_path / (impl + '.pt'))
        env = dict(os.environ, NRT_BWD_IMPL=impl)
        subprocess.run([sys.executable, '-c', code, f], check=True, env=env, timeout=300)
        res[impl] = torch.load(f)
    for name in res['q']:
        a, b = res['q'][name], res['tc'][name]
        s = b.abs().max().item()
        assert s > 0 and torch.isfinite(a).all(), name
        tol = 2e-3 if name.startswith('w') else 2e-5
        err = (a - b).abs().max().item()
        print(f'{name}: max |q - tc| = {err:.3e} of scale {s:.3e}')
        assert err <= tol * s, (name, err, s)
