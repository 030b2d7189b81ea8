"""ERP depth -> distance: oracle and host-side grids against the reference's own ERPDepth2Dist (golden
erp_small.npz, CPU), the fused CUDA kernel against both (GPU), full-size properties."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'erp_small.npz')


def _case(tag):
    g = np.load(GOLD)
    return {k[2:]: torch.from_numpy(g[k]) if g[k].ndim else int(g[k]) for k in g.files if k.startswith(tag + '_')}


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_oracle_matches_reference_golden(tag):
    from oracle.erp_oracle import erp_depth2dist
    c = _case(tag)
    out = erp_depth2dist(c['depth'], c['c2e'], c['coor'], c['rays'], c['s'])
    assert torch.equal(out, c['dist'])


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_host_grids_match_reference_constructor(tag):
    """The three static grids the drop-in builds on the host == the ones the reference's constructor builds."""
    from naruto_b200 import erp
    c = _case(tag)
    s, hw = c['s'], tuple(c['depth'].shape)
    coor = torch.stack([erp.face_sampling_grid(u, v, hw, s) for u, v in zip(erp._FACE_U_DEG, erp._FACE_V_DEG)])
    assert (coor - c['coor']).abs().max() <= 2e-6
    assert torch.equal(erp.texel_rays(s), c['rays'])
    g = erp.cube_to_pano_grid(s, hw)
    assert torch.equal(g[..., 2], c['c2e'][..., 2]), 'face ids'
    assert (g - c['c2e']).abs().max() <= 2e-6


def test_constructor_refuses_cpu():
    from naruto_b200.erp import ERPDepth2Dist
    with pytest.raises(RuntimeError):
        ERPDepth2Dist(16, (16, 32), 'cpu')


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_kernel_matches_reference_golden(tag):
    from naruto_b200.erp import ERPDepth2Dist
    c = _case(tag)
    hw = tuple(c['depth'].shape)
    # fed with the reference's own grids: the texel choices are identical, only fp32 interpolation order can differ
    m = ERPDepth2Dist(c['s'], hw, 'cuda', grids=(c['c2e'], c['coor'], c['rays']))
    out = m(c['depth'].cuda().reshape(1, 1, *hw)).cpu()
    ref = c['dist']
    assert out.shape == ref.shape
    assert ((out - ref).abs() <= 2e-6 * ref.abs()).all(), ((out - ref).abs() / ref.abs()).max()
    # with its own host-built grids: same result wherever the nearest-texel choice agrees (a borderline coordinate may pick
    # the neighbouring texel when a grid value differs in the last bit)
    m2 = ERPDepth2Dist(c['s'], hw, 'cuda')
    out2 = m2(c['depth'].cuda()).cpu()
    frac = ((out2 - ref).abs() <= 2e-6 * ref.abs()).float().mean().item()
    assert frac >= 0.995, frac


@pytest.mark.gpu
def test_kernel_matches_oracle_other_size():
    from naruto_b200.erp import ERPDepth2Dist
    from oracle.erp_oracle import erp_depth2dist
    s, hw = 48, (56, 120)
    m = ERPDepth2Dist(s, hw, 'cuda')
    g = torch.Generator().manual_seed(3)
    depth = 0.5 + 4 * torch.rand(*hw, generator=g)
    depth[torch.rand(*hw, generator=g) < 0.05] = 1e8
    out = m(depth.cuda()).cpu()
    ref = erp_depth2dist(depth, m.c2e_grid.cpu(), m.face_coor.cpu(), m.face_rays.cpu(), s)
    assert ((out - ref).abs() <= 2e-6 * ref.abs()).all()


@pytest.mark.gpu
def test_full_size_properties():
    """1024 x 2048 panorama, 512 skybox (the simulator's shape): a constant depth d maps to d * |ray| with
    1 <= |ray| <= sqrt(3); distances scale linearly with depth; zero stays zero."""
    from naruto_b200.erp import ERPDepth2Dist
    m = ERPDepth2Dist(512, (1024, 2048), 'cuda')
    one = m(torch.ones(1, 1, 1024, 2048, device='cuda'))
    assert one.shape == (1024, 2048)
    assert one.min() >= 1.0 - 1e-6 and one.max() <= 3 ** 0.5 + 1e-5
    d = torch.rand(1024, 2048, device='cuda') + 0.5
    a, b = m(d), m(2 * d)
    assert torch.allclose(b, 2 * a, rtol=1e-6, atol=0)
    assert m(torch.zeros(1024, 2048, device='cuda')).abs().max() == 0
