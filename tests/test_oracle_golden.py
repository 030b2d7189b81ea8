"""CPU: the oracle restatement (oracle/naruto_oracle.py) against golden vectors produced by the
reference's own Python (oracle/make_golden.py), plus the known-answer anchors of SURVEY.md 8c."""
import math

import numpy as np
import pytest
import torch

from conftest import golden_params, load_golden, t
from oracle import naruto_oracle as no
from oracle import tcnn_shim


def test_anchor_level_table(spec):
    # resolutions, entries per level, total parameter count at office0.
    # Caveat at the unpinned tcnn boundary: the shim (and csrc/api.cu) evaluate grid_scale in fp64 and round once, real tcnn in
    # fp32 (exp2f(level * log2f(s)) * base - 1).  The top level's scale is exactly 274.0 in real arithmetic, so tcnn's
    # ceil() + 1 can land on 275 or 276 there; the level is hashed, so its resolution only enters through the hash of the
    # cell coordinates (scale-driven), not through a stride -- the features are unaffected either way.
    assert [lv['res'] for lv in spec.table] == [16, 20, 24, 29, 35, 42, 50, 61, 73, 89, 107, 129, 156, 189, 228, 275]
    assert [lv['size'] for lv in spec.table] == [4096, 8000, 13824, 24392, 42880] + [65536] * 11
    assert spec.n_grid_entries * spec.n_features == 1628176
    assert abs(spec.per_level_scale - 1.208780691248879) < 1e-12
    assert spec.uncert_dims == [49, 56, 35]
    assert spec.n_samples == 43


def test_anchor_oneblob_rows_sum_to_one(spec):
    x = torch.rand(1000, 3) * 1.4 - 0.2
    ob = no.oneblob_features(x, spec).reshape(1000, 3, 16)
    assert torch.allclose(ob.sum(-1), torch.ones(1000, 3), atol=2e-6)


def test_anchor_all_positive_sdf_truncates_at_first_sample(spec):
    # no sign change -> argmax 0 -> only z < z[0] + trunc survive
    z = torch.linspace(0, 5, 43)[None]
    w = no.sdf_to_weights(torch.full((1, 43), 0.3), z, spec)
    assert (w[0, z[0] >= 0.1] == 0).all() and abs(w.sum().item() - 1) < 1e-5


def test_anchor_initial_uncertainty(spec):
    # grid initialised to 3 -> uncert_map = sum w^2 (softplus(3)+0.01)
    raw = torch.zeros(1, 43, 5)
    raw[..., 3] = torch.linspace(1, -1, 43)
    raw[..., 4] = 3.0
    z = torch.linspace(0, 5, 43)[None]
    out = no.composite(raw, z, spec)
    exp = (out['weights'] ** 2).sum() * (math.log1p(math.exp(3.0)) + 0.01)
    assert abs(out['uncert_map'].item() - exp.item()) < 1e-6
    assert abs(out['acc_map'].item() - 1) < 1e-5
    assert abs(out['disp_map'].item() - 1 / out['depth'].item()) < 1e-4


def test_uncert_lookup_manual_matches_grid_sample(spec):
    g = torch.Generator().manual_seed(5)
    grid = torch.rand(spec.uncert_dims, generator=g)
    x = torch.rand(4000, 3, generator=g) * 1.3 - 0.15
    a = no.uncert_lookup(x, grid)
    b = no.uncert_lookup_manual(x, grid)
    assert torch.allclose(a, b, atol=2e-6)


@pytest.mark.parametrize('tag', ['small', 'wide'])
def test_points_match_reference(spec, tag):
    g = load_golden(f'points_{tag}')
    P = golden_params(g, spec)
    x = t(g['x'])
    assert np.array_equal(no.hash_features(x, P.grid, spec).numpy(), g['hash'])
    assert np.array_equal(no.oneblob_features(x, spec).numpy(), g['oneblob'])
    assert np.array_equal(no.uncert_lookup(x, P.uncert_grid).numpy(), g['uncert'])
    np.testing.assert_allclose(no.decode(x, P, spec).numpy(), g['raw'], rtol=1e-6, atol=1e-7)
    su, geo = no.query_sdf(x[None], P, spec, return_geo=True, return_uncert=True)
    np.testing.assert_allclose(su[0].numpy(), g['sdf_uncert'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(geo[0].numpy(), g['geo'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(no.query_sdf(x[None], P, spec)[0].numpy(), g['sdf'], rtol=1e-6, atol=1e-7)
    assert np.array_equal(no.query_sdf(x[None], P, spec, embed=True)[0].numpy(), g['embed'])


@pytest.mark.parametrize('tag', ['small', 'wide'])
@pytest.mark.parametrize('perturb', [0, 1])
def test_render_matches_reference(spec, tag, perturb):
    g = load_golden(f'render_{tag}_p{perturb}')
    P = golden_params(g, spec)
    u = t(g['u']) if perturb else None
    r = no.render_rays(t(g['rays_o']), t(g['rays_d']), t(g['target_d']), P, spec, u=u)
    assert np.array_equal(r['z_vals'].numpy(), g['z_vals'])
    np.testing.assert_allclose(r['raw'].numpy(), g['raw'], rtol=1e-6, atol=1e-7)
    for k in ('rgb', 'depth', 'disp_map', 'acc_map', 'depth_var', 'uncert_map'):
        np.testing.assert_allclose(r[k].numpy(), g[k], rtol=2e-6, atol=1e-7, err_msg=k)


@pytest.mark.parametrize('tag', ['small', 'wide'])
def test_train_losses_and_grads_match_reference(spec, tag):
    from oracle.make_golden import grad_probe_idx
    g = load_golden(f'train_{tag}')
    P = golden_params(g, spec).clone(requires_grad=True)
    ret = no.forward_train(t(g['rays_o']), t(g['rays_d']), t(g['target_rgb']), t(g['target_d']), P, spec, u=t(g['u']))
    for k in ('rgb_loss', 'depth_loss', 'sdf_loss', 'fs_loss', 'uncert_loss', 'psnr'):
        np.testing.assert_allclose(ret[k].detach().numpy(), g[k], rtol=2e-6, err_msg=k)
    loss = no.total_loss(ret, spec)
    np.testing.assert_allclose(loss.item(), g['loss'], rtol=2e-6)
    loss.backward()
    for n in ('w1', 'w2', 'w3', 'w4', 'uncert_grid'):
        ref = g[n + '_grad']
        np.testing.assert_allclose(getattr(P, n).grad.numpy(), ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max(), err_msg=n)
    gg = P.grid.grad
    idx = grad_probe_idx(gg.numel())
    np.testing.assert_allclose(gg[idx].numpy(), g['grid_grad_probe'], rtol=1e-4, atol=1e-6 * np.abs(g['grid_grad_probe']).max())
    assert abs(gg.double().norm().item() - float(g['grid_grad_l2'])) <= 1e-5 * float(g['grid_grad_l2'])
    assert int((gg != 0).sum()) == int(g['grid_grad_nnz'])


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_map_volumes_match_reference(spec, tag):
    """oracle.map_volumes vs the reference's own get_map_volumes + query_sdf (oracle/make_golden_volumes.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'map_volumes_small.npz'))
    P = no.init_params(spec, seed=int(g[f'seed_{tag}']), grid_range=float(g[f'range_{tag}']), uncert_jitter=1.0)
    um, sdf = no.map_volumes(P, spec, 0.25)
    assert tuple(sdf.shape) == g[f'sdf_{tag}'].shape == (20, 23, 15)
    assert np.abs(sdf.numpy() - g[f'sdf_{tag}']).max() <= 1e-6 * max(1.0, np.abs(g[f'sdf_{tag}']).max())
    assert np.abs(um.numpy() - g[f'uncert_{tag}']).max() <= 1e-5
    assert (g[f'uncert_{tag}'] > 0).any() and (g[f'uncert_{tag}'] == 0).any()       # both sides of the surface mask
