#!/usr/bin/env python
"""Timing of the "next" rows built around the hot path, at the reference's sizes:
  * ERPDepth2Dist: 1024 x 2048 panorama, 512-texel skybox (src/simulator/habitat_simulator.py:63,143);
  * goal-space uncertainty aggregation: office0 volumes 49 x 56 x 35, goal space 25 x 28 x 3, k = 300 targets
    (configs/default.py:93-96, src/planner/naruto_planner.py:596-735).
  * marching cubes: the 5 cm office0 lattice 97 x 111 x 69 (configs/default.py:152 save_mesh_voxel_size), SDF of a room;
    beside it the reference's own extractor compiled where it lies (oracle/_ref), one host core.
Device time by CUDA events (median of 20 after 5 warm-ups); beside it the torch-op restatement of the reference (the oracle)
on the host cores, one call each.  Prints one JSON line."""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from naruto_b200.erp import ERPDepth2Dist                         # noqa: E402
from naruto_b200.planner_handoff import GoalSpace                 # noqa: E402
from oracle.erp_oracle import erp_depth2dist                      # noqa: E402  (cpu baseline leg)
from oracle.planner_oracle import goal_aggregate                  # noqa: E402
from oracle.make_golden_planner import synth_volumes              # noqa: E402
from oracle import mc_ref                                         # noqa: E402
from naruto_b200.marching_cubes import marching_cubes             # noqa: E402


def dev_ms(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    out = {'cores': os.cpu_count(), 'torch_threads': torch.get_num_threads()}
    # ---- ERP ----
    H, W, s = 1024, 2048, 512
    m = ERPDepth2Dist(s, (H, W), 'cuda')
    d = (torch.rand(H, W) * 4 + 0.5)
    dd = d.cuda()
    ms = dev_ms(lambda: m(dd))
    # algorithmic bytes: panorama read + write, the three static grids read once per call
    byt = 4 * (2 * H * W + 3 * H * W + 2 * 6 * s * s + 3 * s * s)
    t0 = time.perf_counter()
    ref = erp_depth2dist(d, m.c2e_grid.cpu(), m.face_coor.cpu(), m.face_rays.cpu(), s)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    err = ((m(dd).cpu() - ref).abs() / ref.abs()).max().item()
    out['erp_depth2dist'] = {'shape': [H, W], 'skybox': s, 'device_ms': round(ms, 4), 'algorithmic_GB_per_s': round(byt / ms / 1e6, 1),
                             'oracle_cpu_ms': round(cpu_ms, 1), 'max_rel_err_vs_oracle': err}
    # ---- goal-space aggregation ----
    dims = (49, 56, 35)
    uncert, sdf = synth_volumes(dims, seed=52)
    gs = GoalSpace(dims, gs_z_levels=(5, 11, 17), uncert_top_k_subset=300)
    ud, sd = torch.from_numpy(uncert).cuda(), torch.from_numpy(sdf).cuda()
    topk = gs.select_targets(ud)
    ms = dev_ms(lambda: gs.uncertainty_aggregation_v2([ud, sd], topk_vxl=topk, force_running=True))
    ok, res = gs.uncertainty_aggregation_v2([ud, sd], topk_vxl=topk, force_running=True)
    gx, gy, gz = torch.meshgrid(gs.gs_x_range, gs.gs_y_range, gs.gs_z_range, indexing='ij')
    t0 = time.perf_counter()
    coll, aggre = goal_aggregate(torch.from_numpy(uncert), torch.from_numpy(sdf), gs.goal_space_pts.cpu(), (gx, gy, gz), topk.cpu(),
                                 5.0, 20.0, 0.8)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    out['goal_aggregate'] = {'dims': list(dims), 'goal_points': int(gs.goal_space_pts.shape[0]), 'targets': int(topk.shape[0]),
                             'device_ms': round(ms, 4), 'oracle_cpu_ms': round(cpu_ms, 1),
                             'valid_pairs': int((coll != 0).sum()), 'collections_equal_oracle': bool(torch.equal(res['gs_uncert_collections'].cpu(), coll))}
    # ---- marching cubes ----
    import numpy as np
    nx, ny, nz = 97, 111, 69
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing='ij')
    d = np.stack([x - 3.3, nx - 4.6 - x, y - 2.7, ny - 3.9 - y, z - 2.2, nz - 3.4 - z]).min(0)
    d = np.minimum(d, np.sqrt((x - 40.2) ** 2 + (y - 61.7) ** 2) - 6.8)
    vol = np.clip(d + 0.02 * np.random.default_rng(0).standard_normal(d.shape), -2.5, 2.5).astype(np.float32)
    vd = torch.from_numpy(vol).cuda()
    marching_cubes(vd, 0.0, 3.0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        v, f = marching_cubes(vd, 0.0, 3.0)
        ts.append((time.perf_counter() - t0) * 1e3)
    rec = {'dims': [nx, ny, nz], 'vertices': int(len(v)), 'faces': int(len(f)), 'wall_ms_device_plus_host_merge': round(statistics.median(ts), 2)}
    if mc_ref.available():
        t0 = time.perf_counter()
        v0, f0 = mc_ref.marching_cubes(vol, 0.0, 3.0)
        rec['reference_cpu_ms_1_core'] = round((time.perf_counter() - t0) * 1e3, 1)
        rec['equal_to_reference'] = bool(np.array_equal(v, v0) and np.array_equal(f, f0))
    out['marching_cubes'] = rec
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
