"""TEST INFRASTRUCTURE -- CPU restatement (torch/numpy) of the ray-sampling half of NARUTO's mapping iteration
("next" row f1): key-frame database, global / current-frame ray sampling, the pose transform
and the uncertainty-aware ActiveRaySampler.  The reference draws its indices with Python's `random.sample`; here every
function takes the index lists as INPUTS, so the restatement is deterministic and the CUDA path can be checked on the same
draws.  Pinned by tests/golden/sampler_*.npz, produced by oracle/make_golden_sampler.py from the reference's OWN classes
(KeyFrameDatabaseNaruto, ActiveRaySampler, get_camera_rays) executed in the build container.

Only tests/ and bench tooling may import this file.  Paths relative to /root/reference; tp/ = third_parties/coslam/.
"""
import numpy as np
import torch


def camera_rays(H, W, fx, fy, cx, cy):
    """tp/datasets/utils.py:24-57, type='OpenGL': dirs[j, i] = [(i - cx)/fx, -(j - cy)/fy, -1], not normalised."""
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing='xy')
    return torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)


def frame_rays(direction, rgb, depth):
    """src/slam/coslam/coslam.py:290-291 / src/slam/coslam/model/keyframe.py:43-44: [H*W, 7] = (dir 3, rgb 3, depth 1)."""
    return torch.cat([direction, rgb, depth[..., None]], dim=-1).reshape(-1, 7)


def valid_depth_mask(rays, depth_trunc):
    """(depth > 0) & (depth <= depth_trunc)   (src/slam/coslam/coslam.py:319; src/slam/coslam/model/keyframe.py:28)"""
    return (rays[..., -1] > 0.0) & (rays[..., -1] <= depth_trunc)


def keyframe_select(rays, idxs, num_rays_to_save):
    """KeyFrameDatabaseNaruto.sample_single_keyframe_rays + add_keyframe (src/slam/coslam/model/keyframe.py:21-60):
    rows `idxs` of the UNFILTERED frame rays (the indices were drawn from range(num_valid)), doubled
    until at least num_rays_to_save rows exist, then truncated."""
    sel = rays[torch.as_tensor(idxs, dtype=torch.long)]
    if sel.shape[0] == 0:
        return sel
    while sel.shape[0] < num_rays_to_save:
        sel = torch.cat([sel, sel], dim=0)
    return sel[:num_rays_to_save]


def sample_global(kf_rays, frame_ids, idxs, num_rays_to_save):
    """KeyFrameDatabase.sample_global_rays (tp/model/keyframe.py:69-79): kf_rays [num_kf, P, 7], frame_ids [num_kf]."""
    idxs = torch.as_tensor(idxs, dtype=torch.long)
    return kf_rays.reshape(-1, 7)[idxs], frame_ids[idxs // num_rays_to_save]


def num_current(sample_num, num_kf, min_pixels_cur):
    """src/slam/coslam/coslam.py:318"""
    return max(sample_num // num_kf, min_pixels_cur)


def assemble(rays_g, ids_g, current_rays, idx_cur, keyframe_every, poses_all):
    """src/slam/coslam/coslam.py:329-344: concatenate global and current rays, map frame ids to pose rows (current frame ->
    row -1 = the last pose), rotate the camera-frame directions, origins = pose translations."""
    idx_cur = torch.as_tensor(idx_cur, dtype=torch.long)
    rays = torch.cat([rays_g, current_rays[idx_cur]], dim=0)
    ids_all = torch.cat([torch.div(ids_g, keyframe_every, rounding_mode='trunc'), -torch.ones(len(idx_cur))]).to(torch.int64)
    rays_d_cam, target_s, target_d = rays[..., :3], rays[..., 3:6], rays[..., 6:7]
    rays_d = torch.sum(rays_d_cam[..., None, None, :] * poses_all[ids_all, None, :3, :3], -1)
    rays_o = poses_all[ids_all, None, :3, -1].repeat(1, rays_d.shape[1], 1).reshape(-1, 3)
    return rays_o, rays_d.reshape(-1, 3), target_s, target_d


def active_pool_uncertainty(rays_o, rays_d, target_d, n_cur, uncert_vol, bbox, base_sample_num, oversample_mul):
    """First half of ActiveRaySampler.sample_rays (src/slam/coslam/active_ray_sampler.py:105-122): uncertainty looked up at
    the back-projected end point of every pool ray (rows base_sample_num .. N - ceil(n_cur / mul))."""
    pts = rays_o + rays_d * target_d
    pts = pts[base_sample_num:-n_cur // oversample_mul]
    pts_loc = ((pts - torch.tensor(bbox, dtype=torch.float32)[:, 0]) * 10).numpy()
    pts_idx = pts_loc.round().astype(int)
    for a in range(3):
        pts_idx[:, a] = np.clip(pts_idx[:, a], 0, uncert_vol.shape[a] - 1)
    return np.asarray(uncert_vol)[pts_idx[:, 0], pts_idx[:, 1], pts_idx[:, 2]]


def active_select(rays_o, rays_d, target_s, target_d, n_cur, uncert_vol, bbox, base_sample_num=2048, num_uncert_sample=500,
                  oversample_mul=4):
    """ActiveRaySampler.sample_rays (src/slam/coslam/active_ray_sampler.py:77-149).  Returns the four recombined tensors
    [K lowest-uncertainty pool rays | first base-K global rays | last ceil(n_cur/mul) current rays] and the chosen pool
    indices.  np.argpartition leaves the order (and the choice among ties at the K-th value) unspecified."""
    pu = active_pool_uncertainty(rays_o, rays_d, target_d, n_cur, uncert_vol, bbox, base_sample_num, oversample_mul)
    min_indices = np.argpartition(pu, num_uncert_sample, axis=None)[:num_uncert_sample]
    tail = -n_cur // oversample_mul

    def comb(t):
        return torch.cat([t[min_indices + base_sample_num], t[:base_sample_num - num_uncert_sample], t[tail:]])

    return comb(rays_o), comb(rays_d), comb(target_s), comb(target_d), min_indices, pu
