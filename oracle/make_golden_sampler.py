"""TEST INFRASTRUCTURE -- golden vectors for the ray-sampling rows from the reference's OWN classes.

Runs in the build container only (needs /root/reference): imports KeyFrameDatabaseNaruto
(src/slam/coslam/model/keyframe.py), ActiveRaySampler (src/slam/coslam/active_ray_sampler.py) and get_camera_rays
(third_parties/coslam/datasets/utils.py) unmodified, drives them with Python's `random` seeded and `random.sample` wrapped so
that every index list it hands out is recorded, and stores inputs, the recorded draws and outputs in
tests/golden/sampler_small.npz.  (`Tensor.cuda` is patched to the identity: ActiveRaySampler hard-codes `.cuda()`.)

    python -m oracle.make_golden_sampler
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402


def main():
    assert ref_harness.reference_available()
    ref_harness._install_stubs()
    with ref_harness._in_ref_dir():
        from src.slam.coslam.active_ray_sampler import ActiveRaySampler
        from src.slam.coslam.model.keyframe import KeyFrameDatabaseNaruto
        from third_parties.coslam.datasets.utils import get_camera_rays
    torch.Tensor.cuda = lambda self, *a, **k: self          # CPU container

    H, W, fx, fy = 34, 60, 30.0, 30.0
    cx, cy = 29.5 // 1, 16.5 // 1                            # floor-divided intrinsics
    P, num_kf, every = 300, 8, 5
    depth_trunc = 100.0
    cfg = {'cam': {'depth_trunc': depth_trunc}, 'mapping': {'sample': 64, 'min_pixels_cur': 10}}
    bbox = [[-2.2, 2.6], [-3.4, 2.1], [-1.4, 2.0]]
    g = torch.Generator().manual_seed(5)
    direction = get_camera_rays(H, W, fx, fy, cx, cy)        # [H, W, 3]

    draws = []
    real_sample = random.sample

    def recording_sample(pop, k):
        out = real_sample(pop, k)
        draws.append(np.asarray(out, dtype=np.int64))
        return out

    random.sample = recording_sample
    random.seed(11)

    kf = KeyFrameDatabaseNaruto(cfg, H, W, num_kf, P, 'cpu')
    frames, poses = [], []
    for f in range(4):
        rgb = torch.rand(1, H, W, 3, generator=g)
        depth = torch.rand(1, H, W, generator=g) * 3 + 0.5
        depth[torch.rand(1, H, W, generator=g) < 0.15] = 0.0      # missing depth
        if f == 2:
            depth[:, 3:, :] = 0.0                                 # a frame with fewer valid pixels than P (doubling rule)
        q = torch.linalg.qr(torch.randn(3, 3, generator=g))[0]
        c2w = torch.eye(4)
        c2w[:3, :3] = q
        c2w[:3, 3] = torch.rand(3, generator=g) - 0.5
        poses.append(c2w)
        frames.append((rgb, depth))
    out = {'direction': direction.numpy(), 'H': H, 'W': W, 'fx': fx, 'fy': fy, 'cx': cx, 'cy': cy, 'P': P, 'every': every,
           'depth_trunc': depth_trunc, 'bbox': np.asarray(bbox, dtype=np.float32)}
    for f in range(3):                                           # three key frames, the fourth frame is "current"
        rgb, depth = frames[f]
        n0 = len(draws)
        kf.add_keyframe({'direction': direction[None], 'rgb': rgb, 'depth': depth, 'frame_id': f * every}, filter_depth=True)
        out[f'kf{f}_rgb'], out[f'kf{f}_depth'] = rgb.numpy(), depth.numpy()
        out[f'kf{f}_idxs'] = draws[n0]
    out['kf_rays'] = kf.rays[:3].numpy()
    out['kf_frame_ids'] = kf.frame_ids.numpy()

    # one body of global_BA's sampling (src/slam/coslam/coslam.py:302-359) with the reference's own objects
    sampler = ActiveRaySampler(cfg, num_uncert_sample=20, oversample_mul=4)
    sample_num, min_pixels_cur = sampler.oversample_num, sampler.min_pixels_cur           # 256, 40
    rgb, depth = frames[3]
    current_rays = torch.cat([direction[None], rgb, depth[..., None]], dim=-1).reshape(-1, 7)
    n0 = len(draws)
    rays_g, ids_g = kf.sample_global_rays(sample_num)
    out['global_idxs'] = draws[n0]
    num_cur = max(sample_num // len(kf.frame_ids), min_pixels_cur)
    valid = (current_rays[..., -1] > 0.0) & (current_rays[..., -1] <= depth_trunc)
    cur_num_valid = int(valid.sum())
    idx_cur = random.sample(range(0, cur_num_valid), min(cur_num_valid, num_cur))
    out['idx_cur'] = np.asarray(idx_cur, dtype=np.int64)
    out['cur_num_valid'] = cur_num_valid
    out['cur_rgb'], out['cur_depth'] = rgb.numpy(), depth.numpy()
    poses_all = torch.stack(poses[:3] + [poses[3]])
    out['poses_all'] = poses_all.numpy()
    rays = torch.cat([rays_g, current_rays[idx_cur, :]], dim=0)
    ids_all = torch.cat([torch.div(ids_g, every, rounding_mode='trunc'), -torch.ones((len(idx_cur)))]).to(torch.int64)
    rays_d_cam, target_s, target_d = rays[..., :3], rays[..., 3:6], rays[..., 6:7]
    rays_d = torch.sum(rays_d_cam[..., None, None, :] * poses_all[ids_all, None, :3, :3], -1)
    rays_o = poses_all[ids_all, None, :3, -1].repeat(1, rays_d.shape[1], 1).reshape(-1, 3)
    rays_d = rays_d.reshape(-1, 3)
    out['pre_o'], out['pre_d'], out['pre_s'], out['pre_t'] = [t.numpy() for t in (rays_o, rays_d, target_s, target_d)]
    uncert_vol = (torch.rand(49, 56, 35, generator=g) * 3).numpy().astype(np.float32)
    uncert_vol[:, :, :10] = np.round(uncert_vol[:, :, :10])      # ties, as a real lookup volume has
    out['uncert_vol'] = uncert_vol
    ao, ad, as_, at = sampler.sample_rays(rays_o, rays_d, target_s, target_d, idx_cur, uncert_vol, bbox)
    out['act_o'], out['act_d'], out['act_s'], out['act_t'] = [t.numpy() for t in (ao, ad, as_, at)]
    random.sample = real_sample
    path = os.path.join(ROOT, 'tests', 'golden', 'sampler_small.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: getattr(v, 'shape', v) for k, v in out.items()})


if __name__ == '__main__':
    main()
