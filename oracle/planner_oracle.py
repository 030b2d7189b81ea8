"""TEST INFRASTRUCTURE -- CPU restatement of NarutoPlanner.uncertainty_aggregation_v2
(src/planner/naruto_planner.py:596-735) in torch ops, minus logging.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline may import this.  Pinned to the reference's own method by tests/golden/planner_small.npz
(oracle/make_golden_planner.py)."""
import torch


def goal_aggregate(uncert, sdf, goal_pts, gs_xyz, topk_vxl, min_d, max_d, safe_sdf):
    """uncert, sdf: [X,Y,Z] fp32; goal_pts: [G,3] fp32 voxel coords; gs_xyz: three int64 index tensors (any shape with G
    elements); topk_vxl: [k,3] fp32.  Returns (collections [G,k], aggre [G])."""
    Nx, Ny, Nz = sdf.shape
    k = topk_vxl.shape[0]
    gx, gy, gz = gs_xyz
    gp = goal_pts[:, None, :].repeat([1, k, 1])                        # :639
    view = gp - topk_vxl                                               # :640
    dist = torch.norm(view, dim=2)                                     # :641
    valid = (dist < max_d) * (dist > min_d)                            # :644
    unsafe = (gx < 1) + (gx + 1 >= Nx) + (gy < 1) + (gy + 1 >= Ny) + (gz < 1) + (gz + 1 >= Nz)      # :658-660
    unsafe = unsafe + (sdf[gx, gy, gz] < safe_sdf) + (sdf[(gx + 1).clamp(0, Nx - 1), gy, gz] < safe_sdf) + \
        (sdf[(gx - 1).clamp(0, Nx - 1), gy, gz] < safe_sdf)
    unsafe = unsafe + (sdf[gx, (gy + 1).clamp(0, Ny - 1), gz] < safe_sdf) + (sdf[gx, (gy - 1).clamp(0, Ny - 1), gz] < safe_sdf)
    unsafe = unsafe + (sdf[gx, gy, (gz + 1).clamp(0, Nz - 1)] < safe_sdf) + (sdf[gx, gy, (gz - 1).clamp(0, Nz - 1)] < safe_sdf)
    valid = valid.clone()
    valid[unsafe.reshape(-1), :] = False                               # :670
    near = view[valid]                                                 # :675
    t = torch.linspace(0, 1, 30)
    pts = (gp[valid][..., None] - t * near[..., None]).permute(0, 2, 1).long()     # :677-678
    vis = sdf[pts[:, :, 0], pts[:, :, 1], pts[:, :, 2]].min(dim=1)[0] > 0            # :679-682
    valid = valid.masked_scatter(valid.clone(), vis)                   # :691
    tk = topk_vxl.long()
    ku = uncert[tk[:, 0], tk[:, 1], tk[:, 2]][None, :].repeat(goal_pts.shape[0], 1)  # :705-706
    coll = torch.zeros_like(ku)
    coll[valid] = ku[valid]                                            # :707-708
    return coll, coll.sum(dim=1)
