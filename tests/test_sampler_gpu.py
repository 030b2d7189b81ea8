"""GPU: the device-resident ray sampling (naruto_b200/ray_sampler.py -> csrc/sampler.cu through the C-ABI) against the
golden vectors of the reference's own classes and the oracle, on the reference's recorded index draws; plus properties of
the on-device index generator and a full-size end-to-end batch."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as so

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'sampler_small.npz'))
DEV = 'cuda:0'


def T(k):
    return torch.from_numpy(G[k])


def _cfg(sample=64, min_cur=10, every=5, filter_depth=True):
    return {'cam': {'depth_trunc': float(G['depth_trunc'])},
            'mapping': {'sample': sample, 'min_pixels_cur': min_cur, 'keyframe_every': every, 'filter_depth': filter_depth}}


def _db():
    from naruto_b200.ray_sampler import DeviceKeyFrameDatabase
    kf = DeviceKeyFrameDatabase(_cfg(), int(G['H']), int(G['W']), 8, int(G['P']), DEV)
    for f in range(3):
        kf.add_keyframe({'direction': T('direction')[None], 'rgb': T(f'kf{f}_rgb'), 'depth': T(f'kf{f}_depth'),
                         'frame_id': f * int(G['every'])}, filter_depth=True, idxs=G[f'kf{f}_idxs'])
    return kf


def test_camera_rays_and_frame_packing():
    from naruto_b200 import ray_sampler as rs
    d = rs.camera_rays(int(G['H']), int(G['W']), float(G['fx']), float(G['fy']), float(G['cx']), float(G['cy']), DEV)
    assert torch.equal(d.cpu(), T('direction'))
    fr = rs.pack_frame(d[None], T('cur_rgb').to(DEV), T('cur_depth').to(DEV))
    assert torch.equal(fr.cpu(), so.frame_rays(T('direction')[None], T('cur_rgb'), T('cur_depth')))
    assert rs.valid_depth_count(fr, float(G['depth_trunc'])).item() == int(G['cur_num_valid'])


def test_keyframe_database_matches_reference():
    kf = _db()
    assert len(kf) == 3 and torch.equal(kf.frame_ids.cpu(), T('kf_frame_ids'))
    assert torch.equal(kf.rays[:3].cpu(), T('kf_rays'))                 # incl. the doubling rule of key frame 2
    rays, ids = kf.sample_global_rays(256, idxs=G['global_idxs'])
    ro, io = so.sample_global(T('kf_rays'), T('kf_frame_ids'), G['global_idxs'], int(G['P']))
    assert torch.equal(rays.cpu(), ro) and torch.equal(ids.cpu(), io)


def test_batch_assembly_and_active_selection_match_reference():
    from naruto_b200 import ray_sampler as rs
    kf = _db()
    cur = rs.pack_frame(T('direction')[None].to(DEV), T('cur_rgb').to(DEV), T('cur_depth').to(DEV))
    poses = T('poses_all').to(DEV)
    o, d, s, t = rs.sample_mapping_batch(kf, cur, poses, _cfg(), None, None, sampler=None, idxs_global=G['global_idxs'],
                                         idx_cur=G['idx_cur'])
    for a, k in ((o, 'pre_o'), (s, 'pre_s'), (t, 'pre_t')):
        assert torch.equal(a.cpu(), T(k)), k
    assert (d.cpu() - T('pre_d')).abs().max() <= 2e-7                   # 3-term sum: association order of the tensor op
    sampler = rs.DeviceActiveRaySampler(_cfg(), num_uncert_sample=20, oversample_mul=4)
    bbox = G['bbox'].tolist()
    ao, ad, as_, at = sampler.sample_rays(T('pre_o').to(DEV), T('pre_d').to(DEV), T('pre_s').to(DEV), T('pre_t').to(DEV),
                                          list(G['idx_cur']), T('uncert_vol').to(DEV), bbox, want_chosen=True)
    K, n = 20, T('act_o').shape[0]
    assert ao.shape[0] == n
    # everything behind the K uncertainty-selected rows is position-exact
    for a, k in ((ao, 'act_o'), (ad, 'act_d'), (as_, 'act_s'), (at, 'act_t')):
        assert torch.equal(a[K:].cpu(), T(k)[K:]), k
    # the K selected rows: np.argpartition leaves order and tie-breaks unspecified -> compare the selected uncertainty values
    pu = so.active_pool_uncertainty(T('pre_o'), T('pre_d'), T('pre_t'), len(G['idx_cur']), G['uncert_vol'], bbox, 64, 4)
    chosen = sampler.last_chosen.cpu().numpy()
    assert len(set(chosen.tolist())) == K and (np.diff(chosen) > 0).all()
    assert np.array_equal(np.sort(pu[chosen]), np.sort(pu)[:K])
    assert torch.equal(ao[:K].cpu(), T('pre_o')[torch.from_numpy(chosen).long() + 64])
    ref_rows = {tuple(r) for r in np.round(G['act_t'][:K], 6).tolist()}
    # rows strictly below the K-th value must coincide with the reference's choice
    kth = np.sort(pu)[K - 1]
    strict = {tuple(np.round(G['pre_t'][64 + i], 6).tolist()) for i in np.nonzero(pu < kth)[0]}
    assert strict <= ref_rows


@pytest.mark.parametrize('n,k', [(1, 1), (7, 7), (1000, 37), (4096, 4096), (3 * 40800, 8192), (816000, 2048)])
def test_index_generator_draws_without_replacement(n, k):
    from naruto_b200 import ray_sampler as rs
    a = rs.sample_indices(n, k, 1234, DEV).cpu().numpy()
    assert a.min() >= 0 and a.max() < n and len(np.unique(a)) == k
    b = rs.sample_indices(n, k, 1235, DEV).cpu().numpy()
    if n > 4 * k:
        assert len(np.intersect1d(a, b)) < k // 2 + 8                   # another seed, another draw
    if k >= 2048 and n >= 4 * k:
        h, _ = np.histogram(a, bins=16, range=(0, n))
        assert h.min() > 0.6 * k / 16 and h.max() < 1.4 * k / 16        # roughly uniform over the population


def test_full_size_batch_on_device():
    """Reference sizes: 40 800 rays per key frame, 8 192 global + current rays, K = 500 of the pool, 680 x 1200 frame."""
    from naruto_b200 import ray_sampler as rs
    H, W, P, n_kf = 680, 1200, 40800, 12
    cfg = {'cam': {'depth_trunc': 100.0}, 'mapping': {'sample': 2048, 'min_pixels_cur': 100, 'keyframe_every': 5, 'filter_depth': True}}
    g = torch.Generator().manual_seed(0)
    direction = rs.camera_rays(H, W, 600.0, 600.0, 599.0, 339.0, DEV)
    kf = rs.DeviceKeyFrameDatabase(cfg, H, W, 64, P, DEV)
    depth = (torch.rand(1, H, W, generator=g) * 4 + 0.3)
    depth[torch.rand(1, H, W, generator=g) < 0.02] = 0.0
    rgb = torch.rand(1, H, W, 3, generator=g)
    for f in range(n_kf):
        kf.add_keyframe({'direction': direction[None], 'rgb': rgb.to(DEV), 'depth': depth.to(DEV), 'frame_id': 5 * f}, filter_depth=True)
    # (indices are drawn from range(num_valid) but applied to the unfiltered frame -- SURVEY Appendix B4 -- so a stored ray
    # may have zero depth, exactly as in the reference)
    assert len(kf) == n_kf and bool((kf.rays[:n_kf, :, 2] == -1.0).all()) and bool((kf.rays[:n_kf, :, 6] >= 0).all())
    cur = rs.pack_frame(direction[None], rgb.to(DEV), depth.to(DEV))
    poses = torch.eye(4).repeat(n_kf + 1, 1, 1).to(DEV)
    poses[:, :3, 3] = torch.rand(n_kf + 1, 3, generator=g).to(DEV) - 0.5
    sampler = rs.DeviceActiveRaySampler(cfg, 500, 4)
    vol = torch.rand(49, 56, 35, generator=g).to(DEV)
    bbox = [[-2.2, 2.6], [-3.4, 2.1], [-1.4, 2.0]]
    o, d, s, t = rs.sample_mapping_batch(kf, cur, poses, cfg, vol, bbox, sampler=sampler, seed=3)
    n_cur = max(8192 // n_kf, 400)
    assert o.shape == (2048 + -(-n_cur // 4), 3) and t.shape == (o.shape[0], 1)
    assert torch.isfinite(o).all() and torch.isfinite(d).all() and bool((t >= 0).all())
    assert bool((d[:, 2] == -1.0).all())                                     # identity rotations: z of the camera ray


def test_keyframe_with_no_valid_depth_leaves_its_slot_untouched():
    """filter_depth with device-drawn indices: a frame without a single valid-depth pixel attaches its id and stores nothing,
    like the reference (src/slam/coslam/model/keyframe.py:47-52: rays.shape[1] == 0 -> return)."""
    from naruto_b200.configs import replica_office0
    from naruto_b200.ray_sampler import DeviceKeyFrameDatabase
    cfg = replica_office0()
    H, W, P = 24, 32, 40
    db = DeviceKeyFrameDatabase(cfg, H, W, 4, P, 'cuda')
    g = torch.Generator().manual_seed(0)
    good = {'frame_id': torch.tensor([0]), 'direction': torch.rand(1, H, W, 3, generator=g), 'rgb': torch.rand(1, H, W, 3, generator=g),
            'depth': torch.rand(1, H, W, generator=g) + 0.5}
    bad = dict(good, frame_id=torch.tensor([5]), depth=torch.zeros(1, H, W))
    db.add_keyframe(good, filter_depth=True)
    db.add_keyframe(bad, filter_depth=True)
    torch.cuda.synchronize()
    assert len(db) == 2 and db.frame_ids.tolist() == [0, 5]
    assert db.rays[0].abs().sum() > 0 and (db.rays[0, :, 6] > 0).all()
    assert db.rays[1].abs().sum() == 0, 'no pixel of an all-invalid frame is stored'


def test_direction_image_is_uploaded_once_and_re_uploaded_when_it_changes():
    """batch['direction'] is the same host tensor for every frame of a run (src/slam/coslam/coslam.py:565): the device copy is
    reused while storage, version and a content fingerprint match, and refreshed as soon as the host tensor is written to."""
    from naruto_b200.ray_sampler import device_directions
    from naruto_b200.synthetic import camera_rays
    host = camera_rays(60, 80).unsqueeze(0)
    a = device_directions(host, 'cuda')
    b = device_directions(host.unsqueeze(0).squeeze(0), 'cuda')          # another view object of the same storage
    assert a.data_ptr() == b.data_ptr() and torch.equal(a.cpu(), host)
    host.mul_(2.0)                                                        # in-place write: version counter and content change
    c = device_directions(host, 'cuda')
    assert torch.equal(c.cpu(), host) and not torch.equal(c, a)
    other = host.clone()
    d = device_directions(other, 'cuda')
    assert d.data_ptr() != c.data_ptr() and torch.equal(d.cpu(), other)
