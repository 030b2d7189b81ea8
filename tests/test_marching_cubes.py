"""Iso-surface extraction.  CPU: the numpy oracle against the reference's own marching cubes (golden
mc_small.npz, and the compiled reference itself when oracle/_ref is present).  GPU: the CUDA + host path against golden,
oracle and reference -- vertices and faces must be EQUAL (index work: bit-exact)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'mc_small.npz')
CASES = ['sphere', 'room', 'steep', 'flat']


def _case(name):
    g = np.load(GOLD)
    return g[f'{name}_vol'], float(g[f'{name}_iso']), float(g[f'{name}_trunc']), g[f'{name}_verts'], g[f'{name}_faces']


def _random_volume(seed, shape):
    rng = np.random.default_rng(seed)
    x, y, z = np.meshgrid(*[np.arange(s) for s in shape], indexing='ij')
    c = [s * (0.35 + 0.3 * rng.random()) for s in shape]
    v = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - min(shape) * 0.3
    v = v + 0.1 * rng.standard_normal(shape)
    v[rng.random(shape) < 0.01] = 7.0
    return v.astype(np.float32)


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_golden(name):
    from oracle import mc_oracle
    vol, iso, trunc, verts, faces = _case(name)
    v, f = mc_oracle.marching_cubes(vol, iso, trunc)
    assert v.dtype == np.float64 and f.dtype == np.uint64
    assert np.array_equal(v, verts) and np.array_equal(f, faces)


def test_oracle_matches_compiled_reference_on_random_volumes():
    from oracle import mc_oracle, mc_ref
    if not mc_ref.available():
        pytest.skip('oracle/_ref/libmc_ref.so not built (needs /root/reference)')
    for seed, shape in ((1, (14, 11, 17)), (2, (9, 22, 10)), (3, (16, 16, 16))):
        vol = _random_volume(seed, shape)
        v0, f0 = mc_ref.marching_cubes(vol, 0.05, 3.0)
        v1, f1 = mc_oracle.marching_cubes(vol, 0.05, 3.0)
        assert len(f0) > 50
        assert np.array_equal(v0, v1) and np.array_equal(f0, f1)


def test_triangle_table_is_a_valid_marching_cubes_table():
    """Every listed edge of a case is cut (its two corners on different sides) and cases c / 255-c use the same edges."""
    from oracle.mc_oracle import EDGE_ENDS, tri_table
    t = tri_table()
    assert len(t) == 256 and t[0] == [] and t[255] == []
    for c, row in enumerate(t):
        assert len(row) % 3 == 0 and len(row) <= 15
        for e in row:
            a, b = EDGE_ENDS[e]
            assert ((c >> a) ^ (c >> b)) & 1, (c, e)
        if c not in (85, 170):                          # never emitted by the reference (edge mask 255), stored empty
            cut = {e for e, (a, b) in enumerate(EDGE_ENDS) if ((c >> a) ^ (c >> b)) & 1}
            assert set(row) == cut, c


def test_committed_triangle_table_is_what_the_reference_triangulates_with():
    """csrc/mc_tables.inc against the compiled reference probed one cube case at a time (oracle/probe_mc_tables.py)."""
    from oracle import mc_ref
    if not mc_ref.available():
        pytest.skip('oracle/_ref/libmc_ref.so not built (needs /root/reference)')
    from oracle.mc_oracle import tri_table
    from oracle.probe_mc_tables import probe
    assert tri_table() == probe()


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_kernel_matches_reference_golden(name):
    from naruto_b200.marching_cubes import marching_cubes
    vol, iso, trunc, verts, faces = _case(name)
    v, f = marching_cubes(vol, iso, trunc)
    assert v.dtype == np.float64 and f.dtype == np.uint64 and v.shape[1:] == (3,) and f.shape[1:] == (3,)
    assert np.array_equal(v, verts), 'vertices'
    assert np.array_equal(f, faces), 'faces'


@pytest.mark.gpu
def test_kernel_matches_oracle_and_reference_other_sizes():
    import torch
    from naruto_b200.marching_cubes import marching_cubes
    from oracle import mc_oracle, mc_ref
    for seed, shape in ((11, (33, 20, 27)), (12, (8, 40, 13)), (13, (2, 2, 2)), (14, (1, 5, 5)), (15, (24, 24, 24))):
        vol = _random_volume(seed, shape)
        v1, f1 = mc_oracle.marching_cubes(vol, 0.0, 3.0)
        v, f = marching_cubes(torch.from_numpy(vol).cuda(), 0.0, 3.0)            # device-resident volume
        assert np.array_equal(v, v1) and np.array_equal(f, f1), shape
        if mc_ref.available():
            v0, f0 = mc_ref.marching_cubes(vol, 0.0, 3.0)
            assert np.array_equal(v, v0) and np.array_equal(f, f0), shape


@pytest.mark.gpu
def test_full_size_properties():
    """Dense-sweep size (5 cm office0 lattice, 97 x 111 x 69): a sphere SDF gives a closed 2-manifold -- every edge shared by
    exactly two faces, Euler characteristic 2 -- with all vertices on the iso-surface to interpolation accuracy."""
    import torch
    from naruto_b200.marching_cubes import marching_cubes
    nx, ny, nz = 97, 111, 69
    x, y, z = torch.meshgrid(torch.arange(nx), torch.arange(ny), torch.arange(nz), indexing='ij')
    vol = (torch.sqrt((x - 48.2) ** 2 + (y - 55.4) ** 2 + (z - 34.3) ** 2) - 25.7).float().clamp(-2.5, 2.5).cuda()
    v, f = marching_cubes(vol, 0.0, 3.0)
    assert len(f) > 10000
    r = np.sqrt(((v - np.array([48.2, 55.4, 34.3])) ** 2).sum(1))
    assert np.abs(r - 25.7).max() < 0.05
    fi = f.astype(np.int64)
    e = np.concatenate([fi[:, [0, 1]], fi[:, [1, 2]], fi[:, [2, 0]]])
    e.sort(axis=1)
    _, counts = np.unique(e, axis=0, return_counts=True)
    assert (counts == 2).all()
    assert len(v) - len(counts) + len(f) == 2
