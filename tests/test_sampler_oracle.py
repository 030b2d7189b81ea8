"""CPU: the ray-sampling oracle (oracle/sampler_oracle.py) against golden vectors produced by the reference's own
KeyFrameDatabaseNaruto / ActiveRaySampler / get_camera_rays with its recorded random.sample draws
(oracle/make_golden_sampler.py -> tests/golden/sampler_small.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as so

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'sampler_small.npz'))


def T(k):
    return torch.from_numpy(G[k])


def test_camera_rays_match_reference():
    d = so.camera_rays(int(G['H']), int(G['W']), float(G['fx']), float(G['fy']), float(G['cx']), float(G['cy']))
    assert torch.equal(d, T('direction'))


def test_keyframe_store_matches_reference():
    P = int(G['P'])
    for f in range(3):
        rays = so.frame_rays(T('direction')[None], T(f'kf{f}_rgb'), T(f'kf{f}_depth'))
        idxs = G[f'kf{f}_idxs']
        n_valid = int(so.valid_depth_mask(rays, float(G['depth_trunc'])).sum())
        assert len(idxs) == min(n_valid, P) and idxs.max() < n_valid          # drawn from range(num_valid) ...
        sel = so.keyframe_select(rays, idxs, P)                               # ... applied to the unfiltered rays (B4)
        assert torch.equal(sel, T('kf_rays')[f])
    assert len(G['kf2_idxs']) < P                                             # the doubling rule was exercised


def test_batch_assembly_matches_reference():
    P, every = int(G['P']), int(G['every'])
    rays_g, ids_g = so.sample_global(T('kf_rays'), T('kf_frame_ids'), G['global_idxs'], P)
    cur = so.frame_rays(T('direction')[None], T('cur_rgb'), T('cur_depth'))
    assert int(so.valid_depth_mask(cur, float(G['depth_trunc'])).sum()) == int(G['cur_num_valid'])
    assert len(G['idx_cur']) == so.num_current(256, 3, 40)
    o, d, s, t = so.assemble(rays_g, ids_g, cur, G['idx_cur'], every, T('poses_all'))
    for a, k in ((o, 'pre_o'), (d, 'pre_d'), (s, 'pre_s'), (t, 'pre_t')):
        assert torch.equal(a, T(k)), k


def test_active_selection_matches_reference():
    o, d, s, t = T('pre_o'), T('pre_d'), T('pre_s'), T('pre_t')
    ao, ad, as_, at, chosen, pu = so.active_select(o, d, s, t, len(G['idx_cur']), G['uncert_vol'], G['bbox'].tolist(),
                                                   base_sample_num=64, num_uncert_sample=20, oversample_mul=4)
    for a, k in ((ao, 'act_o'), (ad, 'act_d'), (as_, 'act_s'), (at, 'act_t')):
        assert torch.equal(a, T(k)), k
    assert ao.shape[0] == 64 + -(-len(G['idx_cur']) // 4)
    # the K chosen rays are K of the lowest-uncertainty pool rays (B3: lowest, not highest)
    assert np.sort(pu[chosen]).max() <= np.sort(pu)[19] + 0
