"""TEST INFRASTRUCTURE -- CPU/torch restatement of NARUTO's neural-implicit mapping hot path.

This is the ORACLE: a self-contained, functional restatement (plain torch ops, any device/dtype) of
what the reference's Python computes on the path SURVEY.md section 8 scopes.  It travels to the GPU
box (the reference tree does not) and is the checker for the CUDA kernels.

Pinning: the reference has no tests or golden vectors.  This file is pinned against the
reference's own Python *executed in the build container* (oracle/make_golden.py imports
/root/reference with oracle/tcnn_shim.py standing in for tinycudann and writes tests/golden/*.npz;
tests/test_oracle_golden.py compares).  The tinycudann arithmetic itself is "PARITY UNPINNED" -- see
oracle/tcnn_shim.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  naruto_b200/ never does.

Each function cites the reference lines it restates (paths relative to /root/reference;
tp/ = third_parties/coslam/).
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn.functional as F

from oracle import tcnn_shim


# ------------------------------------------------------------------------------------------------
# configuration
# ------------------------------------------------------------------------------------------------
@dataclass
class FieldSpec:
    """The numbers JointEncodingNaruto reads out of the nested config dict, frozen."""
    bound: List[List[float]]
    n_levels: int = 16
    n_features: int = 2
    log2_hashmap_size: int = 16
    base_resolution: int = 16
    per_level_scale: float = 0.0
    n_bins: int = 16
    hidden_dim: int = 32
    geo_feat_dim: int = 15
    hidden_dim_color: int = 32
    trunc: float = 0.1
    sc_factor: float = 1.0
    near: float = 0.0
    far: float = 5.0
    depth_trunc: float = 100.0
    n_samples_d: int = 32
    n_range_d: int = 11
    range_d: float = 0.1
    perturb: float = 1.0
    uncert_voxel: float = 0.1
    # loss weights (configs/Replica/replica_coslam.yaml:79-85)
    rgb_weight: float = 5.0
    depth_weight: float = 0.1
    sdf_weight: float = 1000.0
    fs_weight: float = 10.0
    uncert_weight: float = 0.005
    smooth_weight: float = 1e-6
    smooth_pts: int = 32
    smooth_vox: float = 0.1
    smooth_margin: float = 0.05
    table: list = field(default_factory=list)
    n_grid_entries: int = 0

    @property
    def n_samples(self):
        return self.n_samples_d + self.n_range_d

    @property
    def uncert_dims(self):
        # src/slam/coslam/model/scene_rep.py:49-52
        return [round((b[1] - b[0]) / self.uncert_voxel + 0.0005) + 1 for b in self.bound]

    def finish(self):
        self.table, self.n_grid_entries = tcnn_shim.level_table(
            self.n_levels, self.base_resolution, self.per_level_scale, self.log2_hashmap_size)
        return self


def spec_from_config(cfg, bound=None, uncert_voxel=0.1):
    """tp/model/scene_rep.py:18-33 (get_resolution) + tp/model/encodings.py:31-33 (per_level_scale)."""
    bound = cfg['mapping']['bound'] if bound is None else bound
    bound = [[float(a), float(b)] for a, b in bound]
    # reference computes dim_max on a float32 tensor
    bt = torch.tensor(bound, dtype=torch.float32)
    dim_max = (bt[:, 1] - bt[:, 0]).max()
    vs = cfg['grid']['voxel_sdf']
    res_sdf = vs if vs > 10 else int(dim_max / vs)
    n_levels, base = 16, 16
    import numpy as np
    pls = float(np.exp2(np.log2(res_sdf / base) / (n_levels - 1)))
    t, c, d = cfg['training'], cfg['cam'], cfg['decoder']
    return FieldSpec(
        bound=bound, n_levels=n_levels, n_features=2, log2_hashmap_size=cfg['grid']['hash_size'],
        base_resolution=base, per_level_scale=pls, n_bins=cfg['pos']['n_bins'],
        hidden_dim=d['hidden_dim'], geo_feat_dim=d['geo_feat_dim'], hidden_dim_color=d['hidden_dim_color'],
        trunc=t['trunc'], sc_factor=cfg['data']['sc_factor'], near=c['near'], far=c['far'],
        depth_trunc=c['depth_trunc'], n_samples_d=t['n_samples_d'], n_range_d=t['n_range_d'],
        range_d=t['range_d'], perturb=t['perturb'], uncert_voxel=uncert_voxel,
        rgb_weight=t['rgb_weight'], depth_weight=t['depth_weight'], sdf_weight=t['sdf_weight'],
        fs_weight=t['fs_weight'], uncert_weight=t.get('uncert_weight', 0.0),
        smooth_weight=t['smooth_weight'], smooth_pts=t['smooth_pts'], smooth_vox=t['smooth_vox'],
        smooth_margin=t['smooth_margin']).finish()


OFFICE0_BOUND = [[-2.2, 2.6], [-3.4, 2.1], [-1.4, 2.0]]          # configs/Replica/office0/coslam.yaml:3


def office0_spec(n_samples_d=32, log2_hashmap_size=16, bound=None):
    """configs/Replica/replica_coslam.yaml + configs/Replica/office0/coslam.yaml, hard-wired (the GPU box has
    no reference tree to read the yaml from)."""
    import numpy as np
    bound = OFFICE0_BOUND if bound is None else bound
    bt = torch.tensor(bound, dtype=torch.float32)
    res_sdf = int((bt[:, 1] - bt[:, 0]).max() / 0.02)
    pls = float(np.exp2(np.log2(res_sdf / 16) / 15))
    return FieldSpec(bound=[list(map(float, b)) for b in bound], per_level_scale=pls,
                     log2_hashmap_size=log2_hashmap_size, n_samples_d=n_samples_d).finish()


@dataclass
class FieldParams:
    """The trainable tensors, with the reference's names in comments."""
    grid: torch.Tensor          # embed_fn.params                         [n_entries*F]
    w1: torch.Tensor            # decoder.sdf_net.model.0.weight          [hidden, 32+48]
    w2: torch.Tensor            # decoder.sdf_net.model.2.weight          [1+geo, hidden]
    w3: torch.Tensor            # decoder.color_net.model.0.weight        [hidden_c, 48+geo]
    w4: torch.Tensor            # decoder.color_net.model.2.weight        [3, hidden_c]
    uncert_grid: torch.Tensor   # uncert_grid                             [Nx,Ny,Nz]

    def tensors(self):
        return [self.grid, self.w1, self.w2, self.w3, self.w4, self.uncert_grid]

    def to(self, *a, **k):
        return FieldParams(*[t.to(*a, **k) for t in self.tensors()])

    def clone(self, requires_grad=False):
        return FieldParams(*[t.detach().clone().requires_grad_(requires_grad) for t in self.tensors()])


def init_params(spec: FieldSpec, seed=0, grid_range=1e-4, uncert_jitter=0.0, dtype=torch.float32):
    """Synthetic weights: grid U(-r, r) (tcnn default r=1e-4), MLPs torch default
    kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)), uncert grid 3 (+ jitter)."""
    g = torch.Generator().manual_seed(seed)

    def lin(o, i):
        k = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * k

    n_in1 = spec.n_levels * spec.n_features + 3 * spec.n_bins
    n_in3 = 3 * spec.n_bins + spec.geo_feat_dim
    grid = (torch.rand(spec.n_grid_entries * spec.n_features, generator=g) * 2 - 1) * grid_range
    ug = torch.full(spec.uncert_dims, 3.0)
    if uncert_jitter:
        ug = ug + (torch.rand(spec.uncert_dims, generator=g) * 2 - 1) * uncert_jitter
    p = FieldParams(grid, lin(spec.hidden_dim, n_in1), lin(1 + spec.geo_feat_dim, spec.hidden_dim),
                    lin(spec.hidden_dim_color, n_in3), lin(3, spec.hidden_dim_color), ug)
    return p.to(dtype)


# ------------------------------------------------------------------------------------------------
# L0: encodings
# ------------------------------------------------------------------------------------------------
def hash_features(x, grid, spec):
    """embed_fn(x): tcnn HashGrid (tp/model/encodings.py:31-46; semantics oracle/tcnn_shim.py)."""
    return tcnn_shim.hash_encode(x, grid, spec.table, spec.n_features)


def oneblob_features(x, spec):
    """embedpos_fn(x): tcnn OneBlob (tp/model/encodings.py:61-71)."""
    return tcnn_shim.oneblob_encode(x, spec.n_bins)


def uncert_lookup(x, uncert_grid):
    """src/slam/coslam/model/scene_rep.py:61-62: grid_sample of the [Nx,Ny,Nz] volume at (2x-1) with the
    (x,y,z) order left un-permuted (world-x walks the LAST axis), trilinear,
    align_corners=False, zero padding."""
    g = (x * 2 - 1)[None, None, None, ...]
    u = F.grid_sample(uncert_grid[None, None, ...], g, align_corners=False)
    return u.reshape(-1)


def uncert_lookup_manual(x, uncert_grid):
    """The same lookup written out (used to cross-check the grid_sample reading the kernels follow):
    continuous index along axis a is ((g+1)*size_a - 1)/2 with g = 2x-1; corners outside contribute 0."""
    dims = uncert_grid.shape                       # [Nx,Ny,Nz]
    g = x * 2 - 1
    # coordinate d of x addresses volume axis (2-d)
    idx, frac = [], []
    for d in range(3):
        size = dims[2 - d]
        c = ((g[:, d] + 1) * size - 1) / 2
        f0 = torch.floor(c)
        idx.append(f0.to(torch.int64))
        frac.append(c - f0)
    out = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
    for corner in range(8):
        w = torch.ones_like(out)
        ok = torch.ones_like(out, dtype=torch.bool)
        ii = []
        for d in range(3):
            bit = (corner >> d) & 1
            i = idx[d] + bit
            w = w * (frac[d] if bit else (1 - frac[d]))
            size = dims[2 - d]
            ok = ok & (i >= 0) & (i < size)
            ii.append(i.clamp(0, size - 1))
        v = uncert_grid[ii[2], ii[1], ii[0]]
        out = out + torch.where(ok, w * v, torch.zeros_like(v))
    return out


# ------------------------------------------------------------------------------------------------
# L1: decoder
# ------------------------------------------------------------------------------------------------
def sdf_head(x, P: FieldParams, spec):
    """query_sdf body (src/slam/coslam/model/scene_rep.py:109-121) + SDFNetNaruto.forward
    (src/slam/coslam/model/decoder.py:29-41): the uncertainty sample rides along as column 0 and is
    re-appended untouched after the bias-free Linear/ReLU/Linear."""
    feat = hash_features(x, P.grid, spec)
    u = uncert_lookup(x, P.uncert_grid)
    blob = oneblob_features(x, spec)
    h = torch.relu(torch.cat([feat, blob], dim=-1) @ P.w1.t())
    o = h @ P.w2.t()
    return o[:, 0], o[:, 1:], u, blob


def decode(x, P: FieldParams, spec):
    """query_color_sdf (src/slam/coslam/model/scene_rep.py:132-148) -> ColorSDFNet_v2_Naruto.forward
    (src/slam/coslam/model/decoder.py:99-116) -> raw[N,5] = [rgb logits(3), sdf, raw uncertainty]."""
    sdf, geo, u, blob = sdf_head(x, P, spec)
    h = torch.relu(torch.cat([blob, geo], dim=-1) @ P.w3.t())
    rgb = h @ P.w4.t()
    return torch.cat([rgb, sdf[:, None], u[:, None]], dim=-1)


def normalise(pts, spec):
    """tp/model/scene_rep.py:172-173."""
    b = torch.tensor(spec.bound, dtype=pts.dtype, device=pts.device)
    return (pts - b[:, 0]) / (b[:, 1] - b[:, 0])


def query_sdf(x, P, spec, return_geo=False, embed=False, return_uncert=False):
    """JointEncodingNaruto.query_sdf (src/slam/coslam/model/scene_rep.py:98-130); x already normalised."""
    flat = x.reshape(-1, 3)
    if embed:
        return hash_features(flat, P.grid, spec).reshape(*x.shape[:-1], -1)
    sdf, geo, u, _ = sdf_head(flat, P, spec)
    out = sdf.reshape(x.shape[:-1])
    if return_uncert:
        out = torch.stack([out, u.reshape(x.shape[:-1])], -1)
    if return_geo:
        return out, geo.reshape(*x.shape[:-1], -1)
    return out


# ------------------------------------------------------------------------------------------------
# L1: ray sampling and compositing
# ------------------------------------------------------------------------------------------------
def sample_z(target_d, spec, u=None):
    """render_rays depth sampling (src/slam/coslam/model/scene_rep.py:158-180).  `u` is the [B,S]
    uniform draw the reference takes from torch.rand (None => perturb off)."""
    B = target_d.shape[0]
    dt, dev = target_d.dtype, target_d.device
    near_d = torch.linspace(-spec.range_d, spec.range_d, steps=spec.n_range_d).to(dt).to(dev)
    z_near = near_d[None, :].repeat(B, 1) + target_d
    z_near[target_d.squeeze(-1) <= 0] = torch.linspace(spec.near, spec.far, steps=spec.n_range_d).to(dt).to(dev)
    if spec.n_samples_d > 0:
        z_uni = torch.linspace(spec.near, spec.far, spec.n_samples_d)[None, :].repeat(B, 1).to(dt).to(dev)
        z, _ = torch.sort(torch.cat([z_uni, z_near], -1), -1)
    else:
        z = z_near
    if u is not None:
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        hi = torch.cat([mid, z[..., -1:]], -1)
        lo = torch.cat([z[..., :1], mid], -1)
        z = lo + (hi - lo) * u
    return z


def sdf_to_weights(sdf, z, spec):
    """JointEncoding.sdf2weights (tp/model/scene_rep.py:64-84)."""
    tr = spec.trunc
    bell = torch.sigmoid(sdf / tr) * torch.sigmoid(-sdf / tr)
    crossing = (sdf[:, 1:] * sdf[:, :-1] < 0.0).to(sdf.dtype)
    first = torch.argmax(crossing, dim=1, keepdim=True)             # 0 when there is no crossing (B6)
    z_surf = torch.gather(z, 1, first)
    keep = (z < z_surf + spec.sc_factor * tr).to(sdf.dtype)
    w = bell * keep
    return w / (w.sum(-1, keepdim=True) + 1e-8)


def composite(raw, z, spec):
    """JointEncodingNaruto.raw2outputs (src/slam/coslam/model/scene_rep.py:66-96)."""
    rgb = torch.sigmoid(raw[..., :3])
    w = sdf_to_weights(raw[..., 3], z, spec)
    rgb_map = (w[..., None] * rgb).sum(-2)
    depth = (w * z).sum(-1)
    depth_var = (w * (z - depth[:, None]) ** 2).sum(-1)
    acc = w.sum(-1)
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    unc = F.softplus(raw[..., 4]) + 0.01
    uncert_map = (w * w * unc).sum(-1)
    return dict(rgb=rgb_map, depth=depth, disp_map=disp, acc_map=acc, depth_var=depth_var,
                uncert_map=uncert_map, weights=w)


def render_rays(rays_o, rays_d, target_d, P, spec, u=None, z=None):
    """JointEncodingNaruto.render_rays (src/slam/coslam/model/scene_rep.py:150-225), n_importance == 0."""
    if z is None:
        z = sample_z(target_d, spec, u)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    B, S = z.shape
    raw = decode(normalise(pts.reshape(-1, 3), spec), P, spec).reshape(B, S, 5)
    out = composite(raw, z, spec)
    out['z_vals'] = z
    out['raw'] = raw
    return out


# ------------------------------------------------------------------------------------------------
# L1: losses
# ------------------------------------------------------------------------------------------------
def sdf_losses(z, target_d, sdf, truncation):
    """get_masks + get_sdf_loss (tp/model/utils.py:81-148), 'l2'."""
    front = (z < (target_d - truncation)).to(z.dtype)
    back = (z > (target_d + truncation)).to(z.dtype)
    has_d = (target_d > 0.0).to(z.dtype)
    near_surf = (1.0 - front) * (1.0 - back) * has_d
    n_fs = torch.count_nonzero(front)
    n_sdf = torch.count_nonzero(near_surf)
    n = n_fs + n_sdf
    fs = F.mse_loss(sdf * front, torch.ones_like(sdf) * front) * (1.0 - n_fs / n)
    sd = F.mse_loss((z + sdf * truncation) * near_surf, target_d * near_surf) * (1.0 - n_sdf / n)
    return fs, sd


def forward_train(rays_o, rays_d, target_rgb, target_d, P, spec, u=None, z=None):
    """JointEncodingNaruto.forward in train mode (src/slam/coslam/model/scene_rep.py:227-287)."""
    r = render_rays(rays_o, rays_d, target_d, P, spec, u=u, z=z)
    td = target_d.squeeze(-1)
    valid = (td > 0.0) & (td < spec.depth_trunc)
    # B13: the bool rgb_weight swallows rgb_missing, every weight is 1
    rgb_loss = F.mse_loss(r['rgb'], target_rgb)
    psnr = -10.0 * torch.log(rgb_loss) / math.log(10.0)
    depth_loss = F.mse_loss(r['depth'][valid], td[valid])
    fs_loss, sdf_loss = sdf_losses(r['z_vals'], target_d, r['raw'][..., 3], spec.trunc * spec.sc_factor)
    U = r['uncert_map'][valid]
    x, y = r['depth'][valid], td[valid]
    # B12: [V,1] x [V] broadcast -> product of means
    uncert_loss = torch.mean((1 / (2 * (U + 1e-9).unsqueeze(-1))) * ((x - y) ** 2)) + 0.5 * torch.mean(torch.log(U + 1e-9))
    return dict(rgb=r['rgb'], depth=r['depth'], rgb_loss=rgb_loss, depth_loss=depth_loss, sdf_loss=sdf_loss,
                fs_loss=fs_loss, psnr=psnr.reshape(1), uncert_loss=uncert_loss, _render=r)


def smoothness(P, spec, offset_u, jitter_u):
    """CoSLAM.smoothness (tp/coslam.py:245-269).  offset_u = the torch.rand(3) draw, jitter_u = the
    torch.rand((1,1,1,3)) draw."""
    b = torch.tensor(spec.bound, dtype=P.grid.dtype, device=P.grid.device)
    n = spec.smooth_pts
    grid_size = (n - 1) * spec.smooth_vox
    offset_max = b[:, 1] - b[:, 0] - grid_size - 2 * spec.smooth_margin
    offset = offset_u.to(b) * offset_max + spec.smooth_margin
    r = torch.arange(0, n - 1, dtype=torch.long, device=b.device)
    coords = torch.stack(torch.meshgrid(r, r, r, indexing='ij'), dim=-1).to(b.dtype)
    pts = (coords + jitter_u.to(b)) * spec.smooth_vox + b[:, 0] + offset
    x = (pts - b[:, 0]) / (b[:, 1] - b[:, 0])
    f = query_sdf(x, P, spec, embed=True)
    tv = ((f[1:] - f[:-1]) ** 2).sum() + ((f[:, 1:] - f[:, :-1]) ** 2).sum() + ((f[:, :, 1:] - f[:, :, :-1]) ** 2).sum()
    return tv / (n ** 3)


def total_loss(ret, spec, smooth=None):
    """CoSLAMNaruto.get_loss_from_ret (src/slam/coslam/coslam.py:154-174)."""
    loss = (spec.rgb_weight * ret['rgb_loss'] + spec.depth_weight * ret['depth_loss']
            + spec.sdf_weight * ret['sdf_loss'] + spec.fs_weight * ret['fs_loss'])
    if smooth is not None and spec.smooth_weight > 0:
        loss = loss + spec.smooth_weight * smooth
    return loss + spec.uncert_weight * ret['uncert_loss']


# ------------------------------------------------------------------------------------------------
# L2: camera rays and one whole mapping iteration (optimiser included)
# ------------------------------------------------------------------------------------------------
def map_volumes(P, spec, voxel_size):
    """get_map_volumes (src/slam/coslam/coslam_utils.py:58-97) with getVoxels (tp/utils.py:26-50): lattice by torch.linspace
    per axis, normalise, query_sdf(return_uncert=True), uncert = softplus + 0.01 kept only where 0 <= sdf < 0.5.
    (The reference also evaluates and discards an `embed` pass)  -> (uncert_vol, sdf_vol)"""
    bb = torch.tensor(spec.bound, dtype=torch.float32)
    lo, hi = [float(v) for v in bb[:, 0]], [float(v) for v in bb[:, 1]]
    n = [round((hi[i] - lo[i]) / voxel_size + 0.0005) for i in range(3)]
    tx, ty, tz = (torch.linspace(lo[i], hi[i], n[i] + 1) for i in range(3))
    pts = torch.stack(torch.meshgrid(tx, ty, tz, indexing='ij'), -1).to(torch.float32)
    x = (pts - bb[:, 0]) / (bb[:, 1] - bb[:, 0])
    su = query_sdf(x, P, spec, return_uncert=True)
    sdf, uncert = su[..., 0], su[..., 1]
    um = F.softplus(uncert) + 0.01
    um = torch.where((sdf >= 0) & (sdf < 0.5), um, torch.zeros_like(um))
    return um, sdf


def camera_rays(H, W, fx, fy, cx, cy):
    """get_camera_rays (tp/datasets/utils.py:24-57), OpenGL convention, un-normalised."""
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing='xy')
    return torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)


class MappingOptimisers:
    """create_optimizer + init_uncert_grid_optim (src/slam/coslam/coslam.py:409-419, 240-243)."""

    def __init__(self, P: FieldParams, lr_decoder=0.01, lr_embed=0.01):
        self.map = torch.optim.Adam(
            [{'params': [P.w3, P.w4, P.w1, P.w2], 'weight_decay': 1e-6, 'lr': lr_decoder},
             {'params': [P.grid], 'eps': 1e-15, 'lr': lr_embed}], betas=(0.9, 0.99))
        self.uncert = torch.optim.Adam(params=[P.uncert_grid], lr=1)


def mapping_iteration(P, opt: MappingOptimisers, rays_o, rays_d, target_rgb, target_d, spec, it,
                      u=None, smooth_draws=None):
    """One body of the global_BA loop (src/slam/coslam/coslam.py:364-399): forward, loss (smooth=True),
    backward, map optimiser step+zero every iteration, uncertainty-grid step+zero every 5th."""
    ret = forward_train(rays_o, rays_d, target_rgb, target_d, P, spec, u=u)
    sm = smoothness(P, spec, *smooth_draws) if smooth_draws is not None else None
    loss = total_loss(ret, spec, sm)
    loss.backward()
    opt.map.step()
    opt.map.zero_grad()
    if (it + 1) % 5 == 0:
        opt.uncert.step()
        opt.uncert.zero_grad()
    return loss.detach(), ret
