"""TEST INFRASTRUCTURE -- golden vectors for the dense uncertainty/SDF sweep from the reference's OWN
`get_map_volumes` (src/slam/coslam/coslam_utils.py:58-97) driving the reference's own `JointEncodingNaruto.query_sdf`.
Build container only (needs /root/reference).  marching_cubes / matplotlib / trimesh, which coslam_utils imports at module
level for its mesh export, are stubbed (absent from this image; not used by get_map_volumes).

    python -m oracle.make_golden_volumes   ->  tests/golden/map_volumes_small.npz
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import naruto_oracle as no   # noqa: E402
from oracle import ref_harness as rh   # noqa: E402
from oracle.make_golden import load_params_into   # noqa: E402


def main():
    assert rh.reference_available()
    rh._install_stubs()
    for name in ('marching_cubes', 'matplotlib', 'matplotlib.pyplot', 'trimesh'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['trimesh'].Trimesh = object                # only named in a return annotation of the mesh export
    cfg = rh.load_reference_config()
    model = rh.build_reference_model(cfg)
    with rh._in_ref_dir():
        from src.slam.coslam.coslam_utils import get_map_volumes
    spec = no.spec_from_config(cfg)
    out = {}
    for tag, seed, rng in (('a', 21, 0.3), ('b', 22, 0.05)):
        P = no.init_params(spec, seed=seed, grid_range=rng, uncert_jitter=1.0)
        load_params_into(model, P)
        model.eval()
        bb = torch.tensor(cfg['mapping']['bound'], dtype=torch.float32)
        uncert, sdf = get_map_volumes(model.query_sdf, bb, 0.25)              # coarse lattice keeps the fixture small
        out[f'uncert_{tag}'], out[f'sdf_{tag}'] = uncert, sdf
        out[f'seed_{tag}'], out[f'range_{tag}'] = seed, rng
    path = os.path.join(ROOT, 'tests', 'golden', 'map_volumes_small.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: getattr(v, 'shape', v) for k, v in out.items()})


if __name__ == '__main__':
    main()
