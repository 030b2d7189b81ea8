"""GPU: the fused dense sweep (nrt_map_volumes) against the golden volumes of the reference's own get_map_volumes and
against the oracle at the planner's real voxel size."""
import os

import numpy as np
import pytest
import torch

from oracle import naruto_oracle as no

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'map_volumes_small.npz'))


def _model(spec, P):
    from test_cuda_parity import make_model
    return make_model(spec, P, torch.device('cuda:0'))


def _check(unc, sdf, unc_ref, sdf_ref):
    assert unc.shape == unc_ref.shape and sdf.shape == sdf_ref.shape
    assert np.abs(sdf - sdf_ref).max() <= 2e-5 * max(1.0, np.abs(sdf_ref).max())
    # the uncertainty is masked by 0 <= sdf < 0.5: voxels whose sdf sits within rounding of a threshold may flip
    safe = (np.abs(sdf_ref) > 1e-4) & (np.abs(sdf_ref - 0.5) > 1e-4)
    assert safe.mean() > 0.99
    assert np.abs(unc - unc_ref)[safe].max() <= 2e-5 * max(1.0, unc_ref.max())


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_map_volumes_match_reference_golden(spec, tag):
    from naruto_b200.map_volumes import get_map_volumes
    P = no.init_params(spec, seed=int(G[f'seed_{tag}']), grid_range=float(G[f'range_{tag}']), uncert_jitter=1.0)
    m = _model(spec, P)
    unc, sdf = get_map_volumes(m, spec.bound, 0.25)
    _check(unc, sdf, G[f'uncert_{tag}'], G[f'sdf_{tag}'])


def test_map_volumes_planner_voxel_vs_oracle(spec):
    from naruto_b200.map_volumes import get_map_volumes
    P = no.init_params(spec, seed=5, grid_range=0.2, uncert_jitter=1.0)
    m = _model(spec, P)
    unc, sdf = get_map_volumes(m, spec.bound, 0.1, to_numpy=False)
    assert tuple(sdf.shape) == (49, 56, 35) and unc.is_cuda
    um_o, sdf_o = no.map_volumes(P, spec, 0.1)
    _check(unc.cpu().numpy(), sdf.cpu().numpy(), um_o.numpy(), sdf_o.numpy())
    # the sweep and the query_sdf API see the same network
    x = torch.rand(1000, 3, device='cuda')
    su = m.query_sdf(x, return_uncert=True)
    assert torch.isfinite(su).all()
