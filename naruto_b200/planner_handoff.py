"""Planner hand-off on the device: the dense sweep of `naruto_b200.map_volumes` stays in HBM and the
goal-space uncertainty aggregation of the planner reads it there.

`GoalSpace` mirrors what `NarutoPlanner.init_local_planner` sets up (src/planner/naruto_planner.py:118-137) and
`GoalSpace.uncertainty_aggregation_v2` mirrors `NarutoPlanner.uncertainty_aggregation_v2` (`:596-735`): same arguments
(`[uncert_vol, sdf_vol]`, numpy like the reference or device tensors), same return value `(goal_space_valid, outputs)` with
`gs_aggre_uncerts [X,Y,Z]`, `topk_uncert_vxl [k,3]` (long), `gs_uncert_collections [G,k]`.  The reference picks its target
voxels with `np.argpartition(uncert, -top_k)[-top_k_subset:]`, whose choice inside the top-k set is unspecified; here the
draw is an argument (`topk_vxl=`, the recorded draw in the parity tests) or, by default, the `top_k_subset` most uncertain
voxels (ties towards the lower flat index).  One CUDA kernel behind `nrt_goal_aggregate`; no torch fallback."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


class GoalSpace:
    def __init__(self, dims, voxel_size=0.1, gs_z_levels=(5, 11, 17), uncert_top_k=4000, uncert_top_k_subset=300,
                 gs_sensing_range=(0.5, 2.0), safe_sdf=0.8, device='cuda'):
        self.dev = torch.device(device)
        if self.dev.type != 'cuda':
            raise RuntimeError('naruto_b200.planner_handoff runs on a CUDA device only (no CPU fallback)')
        self.lib = L.load()
        self.Nx, self.Ny, self.Nz = (int(d) for d in dims)
        self.voxel_size = voxel_size
        self.top_k, self.top_k_subset = int(uncert_top_k), int(uncert_top_k_subset)
        self.sensing = (float(gs_sensing_range[0]), float(gs_sensing_range[1]))
        self.safe_sdf = float(safe_sdf)
        self.gs_x_range = torch.arange(0, self.Nx, 2)
        self.gs_y_range = torch.arange(0, self.Ny, 2)
        if gs_z_levels is None:
            self.gs_z_range = torch.arange(int(1 / voxel_size), self.Nz, int(1 / voxel_size))
        else:
            self.gs_z_range = torch.tensor(list(gs_z_levels))
        gx, gy, gz = torch.meshgrid(self.gs_x_range, self.gs_y_range, self.gs_z_range, indexing='ij')
        self.goal_space_pts = torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1).float().to(self.dev).contiguous()
        self._dims = (C.c_int32 * 3)(self.Nx, self.Ny, self.Nz)
        self._n_valid = torch.zeros(1, dtype=torch.int32, device=self.dev)

    def _vol(self, v):
        t = torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v
        t = t.to(self.dev, torch.float32).contiguous()
        if tuple(t.shape) != (self.Nx, self.Ny, self.Nz):
            raise ValueError(f'volume shape {tuple(t.shape)} != {(self.Nx, self.Ny, self.Nz)}')
        return t

    def select_targets(self, uncert):
        """The top_k_subset most uncertain voxels as [k,3] float voxel coordinates (deterministic stand-in for the reference's
        argpartition draw).  torch.topk is plumbing here: any subset of the top-k set is a legal draw."""
        k = min(self.top_k_subset, uncert.numel())
        idx = torch.topk(uncert.reshape(-1), k, largest=True, sorted=True).indices
        z = idx % self.Nz
        y = (idx // self.Nz) % self.Ny
        x = idx // (self.Nz * self.Ny)
        return torch.stack([x, y, z], dim=1).float()

    @torch.no_grad()
    def uncertainty_aggregation_v2(self, uncert_sdf_vols, force_running=False, topk_vxl=None):
        uncert, sdf = self._vol(uncert_sdf_vols[0]), self._vol(uncert_sdf_vols[1])
        if topk_vxl is None:
            topk_vxl = self.select_targets(uncert)
        topk = topk_vxl.to(self.dev, torch.float32).contiguous()
        G, k = self.goal_space_pts.shape[0], topk.shape[0]
        coll = torch.empty(G, k, dtype=torch.float32, device=self.dev)
        aggre = torch.empty(G, dtype=torch.float32, device=self.dev)
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        L.check(self.lib.nrt_goal_aggregate(L.ptr(uncert), L.ptr(sdf), self._dims, L.ptr(self.goal_space_pts), G, L.ptr(topk), k,
                                            self.sensing[0] / self.voxel_size, self.sensing[1] / self.voxel_size, self.safe_sdf,
                                            L.ptr(coll), L.ptr(aggre), L.ptr(self._n_valid), stream))
        outputs = {
            'gs_aggre_uncerts': aggre.reshape(self.gs_x_range.shape[0], self.gs_y_range.shape[0], self.gs_z_range.shape[0]),
            'topk_uncert_vxl': topk.long(),
            'gs_uncert_collections': coll,
        }
        if int(self._n_valid.item()) == 0 and not force_running:       # the reference's "invalid goal space"
            return False, {}
        return True, outputs
