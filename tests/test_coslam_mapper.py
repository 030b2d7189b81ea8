"""GPU: the Mapper-level drop-in (naruto_b200/coslam_mapper.py) -- the reference's first_frame_mapping / global_BA bodies on
the fused path -- and the one-parameter-set contract behind it (FusedState.bind_model): the fused iteration, the autograd
path, the model's nn.Parameters / state_dict and the torch optimisers' state all alias the same buffers."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(seed=3, n_samples_d=32):
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.scene_rep import JointEncodingNaruto
    cfg = replica_office0(n_samples_d=n_samples_d)
    torch.manual_seed(seed)
    m = JointEncodingNaruto(cfg, torch.tensor(OFFICE0_BOUND)).cuda()
    with torch.no_grad():
        m.embed_fn.params.mul_(300.0)                # U(-0.03, 0.03): away from the all-zero regime of a fresh table
    map_opt = torch.optim.Adam([{'params': m.decoder.parameters(), 'weight_decay': 1e-6, 'lr': cfg['mapping']['lr_decoder']},
                                {'params': m.embed_fn.parameters(), 'eps': 1e-15, 'lr': cfg['mapping']['lr_embed']}], betas=(0.9, 0.99))
    unc_opt = torch.optim.Adam(params=[m.get_uncert_grid(0.1)], lr=1)
    return cfg, m, map_opt, unc_opt


def _loss(cfg, ret):
    t = cfg['training']
    return (t['rgb_weight'] * ret['rgb_loss'] + t['depth_weight'] * ret['depth_loss'] + t['sdf_weight'] * ret['sdf_loss']
            + t['fs_weight'] * ret['fs_loss'] + t['uncert_weight'] * ret['uncert_loss'])


def test_fused_iterations_equal_the_autograd_path_and_share_one_parameter_set():
    """Three iterations (no smoothness term, uncertainty step on the third) through (a) model.forward -> loss.backward ->
    torch.optim.Adam and (b) MappingStep on a FusedState bound to an identical model, same rays and jitter draws."""
    from naruto_b200.mapper import FusedState, MappingStep
    from naruto_b200.synthetic import SyntheticFrame
    from naruto_b200.configs import OFFICE0_BOUND
    cfg, ma, opt_a, unc_a = _model()
    _, mb, opt_b, unc_b = _model()
    mb.load_state_dict(ma.state_dict())
    B = 160
    state = FusedState(mb.plan, 'cuda').bind_model(mb, opt_b, unc_b)
    assert mb.embed_fn.params.data_ptr() == state.P.grid.data_ptr()           # parameters are views of the flat buffer
    ms = MappingStep(mb.plan, cfg, B, 'cuda', state=state, use_graph=True, smooth=False)
    ms.external_random = True
    frame = SyntheticFrame(OFFICE0_BOUND, seed=5)
    g = torch.Generator().manual_seed(9)
    ma.train()
    for it in range(3):
        o, d, rgb, td = [t.cuda() for t in frame.sample(B)]
        u = torch.rand(B, ma.plan.S, generator=g).cuda()
        opt_a.zero_grad()
        if it == 0:
            unc_a.zero_grad()
        ret = ma.forward(o, d, rgb, td, u=u)
        la = _loss(cfg, ret)
        la.backward()
        opt_a.step()
        if it == 2:
            unc_a.step()
            unc_a.zero_grad()
        ms.u.copy_(u)
        ms.step(o, d, rgb, td, with_uncert_step=(it == 2), smooth=False)
        torch.cuda.synchronize()
        lb = (ms.losses[:5] * ms.loss_grad).sum().item()
        assert abs(lb - la.item()) <= 2e-4 * abs(la.item()), (it, lb, la.item())
    state.sync_optimizers()
    for name, pa, pb, tol, frac in (('grid', ma.embed_fn.params, mb.embed_fn.params, 2e-3, 0.995),
                                    ('w1', ma.decoder.sdf_net.model[0].weight, mb.decoder.sdf_net.model[0].weight, 2e-3, 0.99),
                                    ('w3', ma.decoder.color_net.model[0].weight, mb.decoder.color_net.model[0].weight, 2e-3, 0.99),
                                    ('uncert', ma.uncert_grid, mb.uncert_grid, 2e-2, 0.99)):
        dlt = (pa.detach() - pb.detach()).abs().reshape(-1)
        ok = (dlt <= tol).float().mean().item()
        assert ok >= frac, f'{name}: {ok:.4%} within {tol} (max {dlt.max().item():.2e})'
    # the model of path (b) was never touched by a torch optimiser, yet it IS trained: state_dict, optimiser state
    sd = mb.state_dict()
    assert torch.equal(sd['embed_fn.params'], state.P.grid) and not torch.equal(sd['embed_fn.params'], ma.embed_fn.params * 0)
    st = opt_b.state[mb.embed_fn.params]
    assert float(st['step']) == 3 and st['exp_avg'].data_ptr() == state.M.grid.data_ptr()
    assert float(unc_b.state[mb.uncert_grid]['step']) == 1
    assert (mb.uncert_grid.detach() - 3.0).abs().max() > 0.1
    # and the autograd path keeps working on the bound model, on the same numbers (torch Adam continues from step 3)
    o, d, rgb, td = [t.cuda() for t in frame.sample(B)]
    mb.train()
    opt_b.zero_grad()
    before = state.P.w1.clone()
    _loss(cfg, mb.forward(o, d, rgb, td)).backward()
    opt_b.step()
    assert float(opt_b.state[mb.embed_fn.params]['step']) == 4
    assert not torch.equal(state.P.w1, before), 'torch Adam stepped the fused buffer in place'
    state.adopt_optimizer_steps()
    assert state.n_map_steps == 4 and int(state.map_step.item()) == 4


def _fake_slam(cfg, model, map_opt, unc_opt, H, W, active_ray):
    from naruto_b200.ray_sampler import DeviceActiveRaySampler, DeviceKeyFrameDatabase
    from naruto_b200.synthetic import camera_rays
    cfg['mapping']['active_ray'] = active_ray
    s = types.SimpleNamespace()
    s.config, s.model, s.map_optimizer, s.uncert_optim = cfg, model, map_opt, unc_opt
    s.device = torch.device('cuda')
    s.dataset = types.SimpleNamespace(H=H, W=W)
    s.est_c2w_data, s.est_c2w_data_rel, s.step = {}, {}, 0
    s.info_printer = lambda *a, **k: None
    n_save = int(H * W * cfg['mapping']['n_pixels'])
    s.keyframeDatabase = DeviceKeyFrameDatabase(cfg, H, W, 16, n_save, 'cuda')
    if active_ray:
        s.active_ray_sampler = DeviceActiveRaySampler(config=cfg, num_uncert_sample=500, oversample_mul=4)
        s.cached_uncert = torch.rand(49, 56, 35, device='cuda')
    s.rays_d = camera_rays(H, W).cuda()            # the pinhole directions SyntheticFrame derives its depth from
    return s


def _frame_batch(slam, fid, seed):
    from naruto_b200.configs import OFFICE0_BOUND
    from naruto_b200.synthetic import SyntheticFrame
    H, W = slam.dataset.H, slam.dataset.W
    f = SyntheticFrame(OFFICE0_BOUND, seed=seed, H=H, W=W)
    return {'frame_id': torch.tensor([fid]), 'c2w': f.c2w.unsqueeze(0), 'rgb': f.rgb.reshape(1, H, W, 3),
            'depth': f.depth.reshape(1, H, W), 'direction': slam.rays_d.unsqueeze(0).cpu()}, f


@pytest.mark.parametrize('active_ray', [False, True])
def test_first_frame_mapping_and_global_ba_on_the_fused_path(active_ray):
    """A duck-typed CoSLAMNaruto (the attributes the two methods touch) driven through first_frame_mapping and two global_BA
    calls: side effects as in the reference (poses recorded, key frame added, uncertainty grid stepped), the loss falls, and
    the model the planner queries afterwards is the trained one."""
    from naruto_b200 import coslam_mapper as cm
    cfg, m, map_opt, unc_opt = _model()
    cfg['mapping']['first_iters'] = 30
    cfg['mapping']['iters'] = 10
    cfg['mapping']['sample'] = 1024
    H, W = 120, 160
    slam = _fake_slam(cfg, m, map_opt, unc_opt, H, W, active_ray)
    batch, f0 = _frame_batch(slam, 0, seed=11)
    w0 = m.decoder.sdf_net.model[0].weight.detach().clone()
    sdf0 = m.query_sdf(torch.rand(64, 3, device='cuda')).clone()
    ret, loss = cm.first_frame_mapping(slam, batch, n_iters=cfg['mapping']['first_iters'])
    torch.cuda.synchronize()
    assert set(ret) == {'rgb', 'depth', 'rgb_loss', 'depth_loss', 'sdf_loss', 'fs_loss', 'psnr', 'uncert_loss'}
    assert torch.isfinite(loss) and ret['rgb'].shape == (1024, 3)
    assert torch.equal(slam.est_c2w_data[0].cpu(), f0.c2w) and len(slam.keyframeDatabase) == 1
    assert not torch.equal(m.decoder.sdf_net.model[0].weight.detach(), w0), 'the bound model is the trained one'
    assert float(map_opt.state[m.embed_fn.params]['step']) == 30 and float(unc_opt.state[m.uncert_grid]['step']) == 1
    first_loss = loss.item()
    # frames 5 and 10: global_BA with the key-frame database filling up
    losses = []
    for fid in (5, 10):
        batch, f = _frame_batch(slam, fid, seed=11 + fid)
        slam.est_c2w_data[fid] = f.c2w.cuda()
        out = cm.global_BA(slam, batch, fid)
        slam.keyframeDatabase.add_keyframe(batch, filter_depth=cfg['mapping']['filter_depth'])
        torch.cuda.synchronize()
        assert out is not None and torch.isfinite(out[1])
        losses.append(out[1].item())
    assert float(map_opt.state[m.embed_fn.params]['step']) == 50
    assert float(unc_opt.state[m.uncert_grid]['step']) == 1 + 2 * 2          # every 5th iteration of each 10-iteration call
    assert len(slam.keyframeDatabase) == 3
    assert losses[-1] < 20 * first_loss and all(l == l for l in losses)
    # what get_map_volumes / the planner see next is the trained field
    assert not torch.equal(m.query_sdf(torch.rand(64, 3, device='cuda', generator=None)), sdf0)
    # the checkpoint the reference would write carries the trained numbers and the optimiser moments
    sd = m.state_dict()
    assert torch.equal(sd['decoder.sdf_net.model.0.weight'], slam._nrt_fused_mapper.state.P.w1)
    assert map_opt.state_dict()['state'][0]['exp_avg'].abs().max() > 0


def test_pose_refinement_is_refused_not_ignored():
    from naruto_b200 import coslam_mapper as cm
    from naruto_b200._lib import NrtError
    cfg, m, map_opt, unc_opt = _model()
    cfg['tracking']['disable'] = False
    slam = _fake_slam(cfg, m, map_opt, unc_opt, 60, 80, False)
    for fid in (0, 5):
        batch, f = _frame_batch(slam, fid, seed=fid)
        slam.est_c2w_data[fid] = f.c2w.cuda()
        slam.keyframeDatabase.add_keyframe(batch)
    batch, f = _frame_batch(slam, 10, seed=10)
    slam.est_c2w_data[10] = f.c2w.cuda()
    with pytest.raises(NrtError):
        cm.global_BA(slam, batch, 10)
    with pytest.raises(NrtError):      # and the model itself refuses rays that want a gradient
        m.train()
        o = torch.zeros(8, 3, device='cuda', requires_grad=True)
        m.forward(o, torch.ones(8, 3, device='cuda'), torch.zeros(8, 3, device='cuda'), torch.ones(8, 1, device='cuda'))
