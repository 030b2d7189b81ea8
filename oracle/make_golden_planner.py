"""Writes tests/golden/planner_small.npz from the reference's OWN NarutoPlanner.uncertainty_aggregation_v2
(src/planner/naruto_planner.py:596-735), imported unmodified from /root/reference and run on the CPU (`Tensor.cuda` is
patched to the identity for the duration of the call; absent third-party modules are stubbed).  The
np.argpartition draw of the target voxels is recorded with the outputs.  Build container only:
    python -m oracle.make_golden_planner"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_planner_class():
    from oracle import ref_harness
    ref_harness._install_stubs()
    for name in ['habitat_sim', 'quaternion', 'matplotlib', 'matplotlib.pyplot', 'open3d', 'trimesh', 'marching_cubes', 'pytorch3d',
                 'pytorch3d.transforms']:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    for _ in range(8):
        try:
            from src.planner.naruto_planner import NarutoPlanner
            return NarutoPlanner
        except ModuleNotFoundError as e:
            sys.modules[e.name] = types.ModuleType(e.name)
    raise RuntimeError('cannot import the reference planner')


def synth_volumes(dims, seed):
    """A room: SDF = distance (voxels) to the nearest wall of a box inset in the volume, minus a pillar; uncertainty high in a
    corner region, zero where sdf is far from the surface (as get_map_volumes masks it)."""
    g = torch.Generator().manual_seed(seed)
    X, Y, Z = dims
    x, y, z = torch.meshgrid(torch.arange(X), torch.arange(Y), torch.arange(Z), indexing='ij')
    d = torch.stack([x - 1.5, X - 2.5 - x, y - 1.5, Y - 2.5 - y, z - 0.5, Z - 1.5 - z]).float().min(0)[0]
    pillar = torch.sqrt((x - X * 0.55) ** 2 + (y - Y * 0.45) ** 2) - 2.5
    sdf = torch.minimum(d, pillar) + 0.05 * torch.rand(dims, generator=g)
    uncert = torch.rand(dims, generator=g) * ((sdf >= 0) & (sdf < 5)).float()
    uncert[: X // 3, : Y // 3] *= 3.0
    return uncert.numpy().astype(np.float32), sdf.numpy().astype(np.float32)


def main():
    NarutoPlanner = load_planner_class()
    out = {}
    torch.Tensor.cuda = lambda self, *a, **k: self
    for tag, dims, zl, topk, sub in (('a', (25, 29, 17), [5, 11], 400, 48), ('b', (49, 56, 35), [5, 11, 17], 4000, 300)):
        np.random.seed(5)
        Nx, Ny, Nz = dims
        pl = object.__new__(NarutoPlanner)
        pl.planner_cfg = types.SimpleNamespace(uncert_top_k_subset=sub, uncert_top_k=topk, gs_sensing_range=[0.5, 2], safe_sdf=0.8)
        pl.voxel_size = 0.1
        pl.step = 0
        pl.info_printer = lambda *a, **k: None
        pl.Nx, pl.Ny, pl.Nz = Nx, Ny, Nz
        # goal space exactly as init_local_planner builds it (src/planner/naruto_planner.py:124-137)
        pl.gs_x_range, pl.gs_y_range, pl.gs_z_range = torch.arange(0, Nx, 2), torch.arange(0, Ny, 2), torch.tensor(zl)
        pl.gs_x, pl.gs_y, pl.gs_z = torch.meshgrid(pl.gs_x_range, pl.gs_y_range, pl.gs_z_range, indexing='ij')
        pl.goal_space_pts = torch.cat([pl.gs_x.reshape(-1, 1), pl.gs_y.reshape(-1, 1), pl.gs_z.reshape(-1, 1)], dim=1).float()
        uncert, sdf = synth_volumes(dims, seed=3 + Nx)
        ok, res = pl.uncertainty_aggregation_v2([uncert, sdf], force_running=True)
        assert ok
        out[f'{tag}_uncert'], out[f'{tag}_sdf'] = uncert, sdf
        out[f'{tag}_zlevels'] = np.array(zl)
        out[f'{tag}_topk'] = res['topk_uncert_vxl'].numpy()
        out[f'{tag}_coll'] = res['gs_uncert_collections'].numpy()
        out[f'{tag}_aggre'] = res['gs_aggre_uncerts'].numpy()
        print(tag, 'valid pairs', int((res['gs_uncert_collections'] != 0).sum()), 'of', res['gs_uncert_collections'].numel())
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'planner_small.npz'), **out)


if __name__ == '__main__':
    main()
