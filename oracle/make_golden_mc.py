"""Writes tests/golden/mc_small.npz with the output of the reference's OWN marching cubes (oracle/_ref/libmc_ref.so =
marching_cubes.cpp compiled where it lies, oracle/build_ref.py) on small synthetic SDF volumes.  Build container only:
    python -m oracle.make_golden_mc"""
import os

import numpy as np

from oracle import build_ref, mc_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def volumes():
    rng = np.random.default_rng(7)
    out = {}
    x, y, z = np.meshgrid(np.arange(20), np.arange(23), np.arange(18), indexing='ij')
    v = (np.sqrt((x - 9.3) ** 2 + (y - 10.1) ** 2 + (z - 8.7) ** 2) - 6.2 + 0.05 * rng.standard_normal(x.shape)).astype(np.float32)
    v[3:6, 3:6, 3:6] = 5.0                                    # beyond the truncation: holes in the surface
    out['sphere'] = (v, 0.0, 3.0)
    # a room like the dense SDF sweep: box walls, a pillar, unobserved (-inf / large) regions, a NaN, iso level off zero
    x, y, z = np.meshgrid(np.arange(31), np.arange(26), np.arange(15), indexing='ij')
    d = np.stack([x - 2.4, 27.7 - x, y - 1.6, 23.2 - y, z - 1.3, 12.9 - z]).min(0)
    d = np.minimum(d, np.sqrt((x - 14.2) ** 2 + (y - 11.7) ** 2) - 2.8).astype(np.float32)
    d += (0.02 * rng.standard_normal(d.shape)).astype(np.float32)
    d[20:24, 5:9, :] = -np.inf
    d[5, 5, 5] = np.nan
    d[25:, 20:, 10:] = 40.0
    out['room'] = (d, 0.25, 3.0)
    # steep field: neighbouring corners jump by more than the internal threshold of 10 -> cells rejected
    s = (rng.standard_normal((12, 12, 12)) * 9).astype(np.float32)
    out['steep'] = (s, 0.0, 100.0)
    out['flat'] = (np.ones((6, 5, 4), np.float32), 0.0, 3.0)  # no surface at all
    return out


def main():
    assert build_ref.build() and mc_ref.available()
    out = {}
    for name, (vol, iso, trunc) in volumes().items():
        v, f = mc_ref.marching_cubes(vol, iso, trunc)
        out[f'{name}_vol'], out[f'{name}_iso'], out[f'{name}_trunc'] = vol, np.float64(iso), np.float64(trunc)
        out[f'{name}_verts'], out[f'{name}_faces'] = v, f
        print(name, vol.shape, 'verts', v.shape, 'faces', f.shape)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'mc_small.npz'), **out)


if __name__ == '__main__':
    main()
