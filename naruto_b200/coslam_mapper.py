"""The reference's own Mapper loop on the fused path: drop-in bodies for `CoSLAMNaruto.first_frame_mapping`
(src/slam/coslam/coslam.py:176-226) and `CoSLAMNaruto.global_BA` (src/slam/coslam/coslam.py:246-407).

    from naruto_b200.coslam_mapper import FusedMappingMixin
    class CoSLAMNaruto(FusedMappingMixin, SlamModel, CoSLAM): ...        # or: CoSLAMNaruto.global_BA = FusedMappingMixin.global_BA

Same arguments (`batch` dict with c2w / rgb / depth / direction, `cur_frame_id`), same side effects on `self`
(`est_c2w_data`, `keyframeDatabase`, the model's parameters, the state of `map_optimizer` / `uncert_optim`), but per
iteration the work is: ray sampling on the device (naruto_b200.ray_sampler: no Python `random.sample`, no 22.8 MB boolean
index, no host round trip of the active sampler) followed by ONE CUDA-graph replay of the fused iteration
(naruto_b200.mapper.MappingStep: render + losses + backward + smoothness + Adam).  The model's nn.Parameters and the torch
optimisers' Adam state are views of the fused buffers (FusedState.bind_model), so everything else CoSLAMNaruto does with
`self.model` -- get_map_volumes, save_mesh, save_ckpt, the autograd path -- keeps seeing the trained numbers.

What is NOT served (raises NrtError instead of silently doing something else): pose refinement inside global_BA
(`tracking.disable: False` with >= 2 key frames -- NARUTO ships with tracking disabled and ground-truth poses), and
`mapping.map_accum_step != 1` / `map_wait_step != 0`.
Deviation kept from the device sampler (documented there): with `mapping.filter_depth`, the reference draws
min(n_valid, num_cur) current-frame rays; here num_cur rays are always drawn from the n_valid valid pixels (with repeats if
there are fewer), so the batch size does not depend on a device-side count and no host synchronisation is needed.
"""
import torch

from . import _lib as L
from .field import FieldTensors
from .mapper import FusedState, MappingStep
from .ray_sampler import (DeviceActiveRaySampler, DeviceKeyFrameDatabase, device_directions, pack_frame, sample_indices,
                          sample_mapping_batch)


class FusedMapper:
    """State that lives next to a CoSLAMNaruto instance: the FusedState bound to its model / optimisers and one MappingStep
    (static buffers + CUDA graphs) per batch size met so far."""

    def __init__(self, slam, process_group=None, use_graph=True):
        self.slam = slam
        model = slam.model
        if not hasattr(model, 'plan') or not hasattr(model, '_tensors'):
            raise L.NrtError('FusedMapper needs naruto_b200.scene_rep.JointEncodingNaruto as the scene model '
                             '(swap the import at src/slam/coslam/coslam.py:22)')
        self.plan, self.cfg = model.plan, slam.config
        self.dev = model.embed_fn.params.device
        if self.dev.type != 'cuda':
            raise L.NrtError('FusedMapper needs the model on a CUDA device -- there is no CPU path')
        self.pg, self.use_graph = process_group, use_graph
        self.state = FusedState(self.plan, self.dev)
        self.state.bind_model(model, getattr(slam, 'map_optimizer', None), getattr(slam, 'uncert_optim', None))
        self.steps = {}
        self.uses = {}
        self._seed = 0

    def step_for(self, n_rays, n_iters=0):
        """The MappingStep of this batch size.  While the key-frame database fills up (its first ~20 frames) the batch size
        changes from call to call (the current-frame share is sample // n_keyframes): a size met for the first time runs its
        iterations as plain launches, and only a size that comes back is captured into CUDA graphs -- a capture costs about as
        much as 20 eager iterations."""
        n_rays = int(n_rays)
        if n_rays not in self.steps:
            if len(self.steps) >= 8:                       # early key frames change the batch size a few times; keep memory bounded
                old = next(iter(self.steps))
                self.steps.pop(old).release_graphs()
                self.uses.pop(old, None)
            self.steps[n_rays] = MappingStep(self.plan, self.cfg, n_rays, self.dev, process_group=self.pg,
                                             use_graph=False, state=self.state)
            self.uses[n_rays] = 0
        self.uses[n_rays] += 1
        ms = self.steps[n_rays]
        if self.use_graph and (self.uses[n_rays] >= 2 or n_iters >= 40):      # (first_frame_mapping: 200 iterations of one size)
            ms.use_graph = True
        return ms

    def next_seed(self):
        self._seed += 1
        return self._seed

    def release(self):
        for s in self.steps.values():
            s.release_graphs()
        self.steps.clear()


def _mapper(slam) -> FusedMapper:
    m = getattr(slam, '_nrt_fused_mapper', None)
    if m is None or m.slam is not slam or m.state._bound is None or m.state._bound[0] is not slam.model:
        m = FusedMapper(slam)
        slam._nrt_fused_mapper = m
    m.state.adopt_optimizer_steps()
    return m


def _ret_dict(ms: MappingStep):
    """The training-mode return dictionary of JointEncodingNaruto.forward (src/slam/coslam/model/scene_rep.py:279-287), from
    the fused iteration's buffers (device tensors; no synchronisation)."""
    l = ms.losses
    return {'rgb': ms.out.rgb, 'depth': ms.out.depth, 'rgb_loss': l[L.LOSS_RGB], 'depth_loss': l[L.LOSS_DEPTH],
            'sdf_loss': l[L.LOSS_SDF], 'fs_loss': l[L.LOSS_FS], 'psnr': l[L.LOSS_PSNR:L.LOSS_PSNR + 1],
            'uncert_loss': l[L.LOSS_UNCERT]}


def _loss_tensor(ms: MappingStep, smooth: bool):
    """get_loss_from_ret's scalar (src/slam/coslam/coslam.py:154-174) as a device tensor."""
    tot = (ms.losses[:5] * ms.loss_grad).sum()
    if smooth and ms.smooth_w > 0:
        tot = tot + ms.smooth_w * ms.smooth_loss[0]
    return tot


def first_frame_mapping(slam, batch, n_iters=100, indices=None):
    """src/slam/coslam/coslam.py:176-226.  `indices` (optional, [n_iters, sample] flat pixel ids = the reference's
    select_samples draws) is a parity hook; by default the pixels are drawn on the device."""
    fm = _mapper(slam)
    cfg, dev = slam.config, fm.dev
    slam.info_printer("First frame mapping...", slam.step, slam.__class__.__name__)
    c2w = batch['c2w'][0].to(dev)
    slam.est_c2w_data[0] = c2w
    slam.est_c2w_data_rel[0] = c2w
    slam.model.train()
    H, W, n = slam.dataset.H, slam.dataset.W, int(cfg['mapping']['sample'])
    frame = pack_frame(device_directions(batch['direction'], dev).squeeze(0), batch['rgb'].squeeze(0).to(dev), batch['depth'].squeeze(0).to(dev))
    poses = c2w.reshape(1, 4, 4).float().contiguous()
    ms = fm.step_for(n, n_iters)
    if cfg['decoder']['uncert_grid']:
        ms.zero_uncert_grad()                                  # self.uncert_optim.zero_grad()
    empty = torch.empty(0, dtype=torch.int64, device=dev)
    lib = fm.plan.lib
    for i in range(n_iters):
        if indices is None:
            idx = sample_indices(H * W, n, (fm.next_seed() << 1) | 1, dev)
        else:
            ids = torch.as_tensor(indices[i], dtype=torch.int64, device=dev)
            # the reference decodes a flat id as (id % H, id // H): row-major pixel (id % H) * W + id // H
            idx = (ids % H) * W + torch.div(ids, H, rounding_mode='trunc')
        # rays_o = c2w translation, rays_d = R d_cam for the n drawn pixels of this one frame (the "current frame" slot)
        L.check(lib.nrt_assemble_rays(None, None, 1, 1, L.ptr(empty), 0, L.ptr(frame), L.ptr(idx), n, L.ptr(poses), 1,
                                      L.ptr(ms.rays_o), L.ptr(ms.rays_d), L.ptr(ms.target_rgb), L.ptr(ms.target_d),
                                      torch.cuda.current_stream().cuda_stream))
        # get_loss_from_ret(ret): no smoothness term here; only the map optimiser steps inside the loop
        ms.step(with_uncert_step=False, smooth=False)
    if cfg['decoder']['uncert_grid']:
        ms.uncert_step()                                       # one uncert_optim.step() on the gradient accumulated over all iterations
    fm.state.sync_optimizers()
    # First frame will always be a keyframe
    slam.keyframeDatabase.add_keyframe(batch, filter_depth=cfg['mapping']['filter_depth'])
    slam.info_printer("First frame mapping done", slam.step, slam.__class__.__name__)
    return _ret_dict(ms), _loss_tensor(ms, smooth=False)


def global_BA(slam, batch, cur_frame_id, draws=None):
    """src/slam/coslam/coslam.py:246-407.  `draws` (optional, per iteration a pair (idxs_global, idx_cur) = the reference's
    random.sample results) is a parity hook."""
    fm = _mapper(slam)
    cfg, dev = slam.config, fm.dev
    kf_every = cfg['mapping']['keyframe_every']
    kfdb = slam.keyframeDatabase
    if not isinstance(kfdb, DeviceKeyFrameDatabase):
        raise L.NrtError('the fused global_BA samples from naruto_b200.ray_sampler.DeviceKeyFrameDatabase '
                         '(swap the import at src/slam/coslam/coslam.py:23)')
    if not (len(kfdb) < 2 or cfg['tracking']['disable']):
        raise L.NrtError('pose refinement inside global_BA (tracking.disable False) is not implemented: the fused renderer has '
                         'no d rays / d pose output; NARUTO ships with tracking disabled')
    # all the KF poses 0, 5, 10, ... then the current frame (poses are fixed: tracking is disabled)
    poses_all = torch.stack([slam.est_c2w_data[i] for i in range(0, cur_frame_id, kf_every)] + [slam.est_c2w_data[cur_frame_id]])
    poses_all = poses_all.to(dev, torch.float32).contiguous()
    slam.model.train()
    current_rays = pack_frame(device_directions(batch['direction'], dev).squeeze(0), batch['rgb'].squeeze(0).to(dev), batch['depth'].squeeze(0).to(dev))
    sampler = getattr(slam, 'active_ray_sampler', None) if cfg['mapping']['active_ray'] else None
    if sampler is not None and not isinstance(sampler, DeviceActiveRaySampler):
        raise L.NrtError('the fused global_BA needs naruto_b200.ray_sampler.DeviceActiveRaySampler as self.active_ray_sampler')
    cached = getattr(slam, 'cached_uncert', None)
    if sampler is not None and cached is not None and not torch.is_tensor(cached):
        cached = torch.as_tensor(cached, dtype=torch.float32, device=dev)
    ms = None
    for i in range(cfg['mapping']['iters']):
        ig, ic = (draws[i] if draws is not None else (None, None))
        o, d, s, t = sample_mapping_batch(kfdb, current_rays, poses_all, cfg, cached, cfg['mapping']['bound'], sampler=sampler,
                                          idxs_global=ig, idx_cur=ic, seed=fm.next_seed())
        if ms is None or ms.B != o.shape[0]:
            ms = fm.step_for(o.shape[0])
            if i == 0 and cfg['decoder']['uncert_grid']:
                ms.zero_uncert_grad()                          # self.uncert_optim.zero_grad() at the top of global_BA
        ms.load_rays(o, d, s, t)
        # loss.backward(); map_optimizer.step(); every 5th iteration of THIS call uncert_optim.step() + zero_grad()
        ms.step(with_uncert_step=cfg['decoder']['uncert_grid'] and (i + 1) % 5 == 0, smooth=True)
    fm.state.sync_optimizers()
    return (_ret_dict(ms), _loss_tensor(ms, smooth=True)) if ms is not None else None


class FusedMappingMixin:
    """Mix into CoSLAMNaruto (before it in the MRO) to route its two mapping entry points through the fused path."""

    def first_frame_mapping(self, batch, n_iters=100):
        return first_frame_mapping(self, batch, n_iters)

    def global_BA(self, batch, cur_frame_id):
        return global_BA(self, batch, cur_frame_id)
