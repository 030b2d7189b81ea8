"""Thin tensor-level wrappers over the C-ABI (naruto_b200/_lib.py): one `FieldPlan` per scene configuration.

Everything here runs on the CUDA device through libnaruto_b200.so; tensors are only containers for device
memory.  Nothing falls back to torch ops or to the CPU.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib as L

ENC_DIMS = 32
OB_DIMS = 48


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


@dataclass
class FieldTensors:
    """The six trainable tensors in the reference's layouts."""
    grid: torch.Tensor
    w1: torch.Tensor
    w2: torch.Tensor
    w3: torch.Tensor
    w4: torch.Tensor
    uncert: torch.Tensor

    def as_list(self):
        return [self.grid, self.w1, self.w2, self.w3, self.w4, self.uncert]

    def c_params(self):
        return L.NrtParams(*[L.ptr(_f32c(t)) for t in self.as_list()])

    def c_grads(self):
        return L.NrtGrads(*[L.ptr(t) for t in self.as_list()])


class RenderBuffers:
    """Outputs of nrt_render_fwd for B rays x S samples."""

    def __init__(self, B, S, device, per_sample=True, weights=False, feat=False):
        f = dict(dtype=torch.float32, device=device)
        self.B, self.S = B, S
        self.rgb = torch.empty(B, 3, **f)
        self.depth = torch.empty(B, **f)
        self.depth_var = torch.empty(B, **f)
        self.acc = torch.empty(B, **f)
        self.disp = torch.empty(B, **f)
        self.uncert = torch.empty(B, **f)
        self.z_vals = torch.empty(B, S, **f) if per_sample else None
        self.raw = torch.empty(B, S, 5, **f) if per_sample else None
        self.weights = torch.empty(B, S, **f) if weights else None
        # saved hash features, tile-major (see NrtRenderOut::feat): padded to whole tiles of 128 points
        self.feat = torch.zeros((B * S + 127) // 128 * 128, ENC_DIMS, **f) if feat else None
        self.masks = torch.zeros((B * S + 127) // 128 * 128, 2, dtype=torch.int32, device=device) if feat else None   # ReLU masks (padded like feat)

    def feat_rows(self):
        """The saved hash features as a plain [B*S, 32] tensor (un-tiles NrtRenderOut::feat; diagnostics and tests)."""
        n = self.B * self.S
        t = self.feat.view(-1, 8, 128, 4).permute(0, 2, 1, 3).reshape(-1, ENC_DIMS)
        return t[:n].contiguous()

    def c_struct(self):
        return L.NrtRenderOut(L.ptr(self.rgb), L.ptr(self.depth), L.ptr(self.depth_var), L.ptr(self.acc), L.ptr(self.disp),
                              L.ptr(self.uncert), L.ptr(self.z_vals), L.ptr(self.raw), L.ptr(self.weights), L.ptr(self.feat),
                              L.ptr(self.masks))


class FieldPlan:
    """nrt_plan_create + typed calls.  `cfg` is the reference's nested config dict; `bound` a [3,2] array."""

    def __init__(self, cfg: dict, bound, uncert_voxel: float = 0.1):
        self.lib = L.load()
        b = torch.as_tensor(bound, dtype=torch.float32).cpu()
        self.bound = b
        # tp/model/scene_rep.py:18-28 (get_resolution) and tp/model/encodings.py:31-33
        dim_max = (b[:, 1] - b[:, 0]).max()
        vs = cfg['grid']['voxel_sdf']
        self.resolution_sdf = vs if vs > 10 else int(dim_max / vs)
        n_levels, base = 16, 16
        self.per_level_scale = float(np.exp2(np.log2(self.resolution_sdf / base) / (n_levels - 1)))
        # src/slam/coslam/model/scene_rep.py:49-52
        self.uncert_dims = [round((b[i, 1] - b[i, 0]).item() / uncert_voxel + 0.0005) + 1 for i in range(3)]
        t, c, d = cfg['training'], cfg['cam'], cfg['decoder']
        enc = cfg['grid']['enc'].lower()
        if not ('hash' in enc or 'tiled' in enc):
            raise L.NrtError(f"naruto_b200 implements grid.enc='HashGrid' only (got {cfg['grid']['enc']})")
        if 'blob' not in cfg['pos']['enc'].lower():
            raise L.NrtError(f"naruto_b200 implements pos.enc='OneBlob' only (got {cfg['pos']['enc']})")
        if not cfg['grid'].get('oneGrid', True) or d.get('tcnn_network', False) or d.get('pred_uncert', False) \
                or not d.get('uncert_grid', False) or t.get('n_importance', 0) > 0 or not cfg['grid'].get('tcnn_encoding', True):
            raise L.NrtError('naruto_b200 implements the shipped NARUTO configuration: oneGrid, nn.Linear decoders, '
                             'uncert_grid, n_importance=0, tcnn_encoding')
        if float(t.get('rgb_missing', 0.05)) == 0.0:
            # src/slam/coslam/model/scene_rep.py:249-250 writes rgb_missing into a BOOL weight tensor: every non-zero value
            # becomes True (weight 1) -- what the kernels implement -- but 0.0 would really mask rays
            raise L.NrtError('training.rgb_missing == 0 is not implemented (the one value for which the reference\'s bool '
                             'rgb_weight masks rays; shipped configs use 0.05)')
        self.c = L.NrtConfig(
            abi_version=L.NRT_ABI_VERSION, n_levels=n_levels, n_features=2, log2_hashmap_size=int(cfg['grid']['hash_size']),
            base_resolution=base, per_level_scale=self.per_level_scale, n_bins=int(cfg['pos']['n_bins']),
            hidden_dim=int(d['hidden_dim']), geo_feat_dim=int(d['geo_feat_dim']), hidden_dim_color=int(d['hidden_dim_color']),
            bound_min=(L.c_f * 3)(*[float(v) for v in b[:, 0]]), bound_max=(L.c_f * 3)(*[float(v) for v in b[:, 1]]),
            uncert_dims=(C.c_int32 * 3)(*self.uncert_dims), trunc=float(t['trunc']), sc_factor=float(cfg['data']['sc_factor']),
            near_z=float(c['near']), far_z=float(c['far']), depth_trunc=float(c['depth_trunc']),
            n_samples_d=int(t['n_samples_d']), n_range_d=int(t['n_range_d']), range_d=float(t['range_d']))
        h = C.c_void_p()
        L.check(self.lib.nrt_plan_create(C.byref(self.c), C.byref(h)))
        self.h = h
        n_grid, S, enc_d = C.c_int64(), C.c_int32(), C.c_int32()
        L.check(self.lib.nrt_plan_sizes(self.h, C.byref(n_grid), C.byref(S), C.byref(enc_d)))
        self.n_grid_floats, self.S = n_grid.value, S.value
        self.perturb = 1 if t['perturb'] > 0 else 0
        self.white_bkgd = bool(t.get('white_bkgd', False))
        if self.white_bkgd:
            raise L.NrtError('white_bkgd=True is not implemented (False at every shipped config)')

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.nrt_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def levels(self):
        sc = (L.c_f * 16)()
        res, size, off = (C.c_int32 * 16)(), (C.c_int32 * 16)(), (C.c_int32 * 16)()
        L.check(self.lib.nrt_plan_levels(self.h, sc, res, size, off))
        return [dict(scale=float(sc[i]), res=int(res[i]), size=int(size[i]), offset=int(off[i])) for i in range(16)]

    # ---- encodings ---------------------------------------------------------------------------
    def encode_fwd(self, grid, x):
        x = _f32c(x)
        out = torch.empty(x.shape[0], ENC_DIMS, dtype=torch.float32, device=x.device)
        L.check(self.lib.nrt_encode_fwd(self.h, L.ptr(_f32c(grid)), L.ptr(x), x.shape[0], L.ptr(out), _stream()))
        return out

    def encode_bwd(self, grid, x, dout, dgrid=None, want_dx=False):
        x, dout = _f32c(x), _f32c(dout)
        dx = torch.empty_like(x) if want_dx else None
        L.check(self.lib.nrt_encode_bwd(self.h, L.ptr(_f32c(grid)), L.ptr(x), x.shape[0], L.ptr(dout), L.ptr(dgrid), L.ptr(dx),
                                        _stream()))
        return dx

    def oneblob_fwd(self, x):
        x = _f32c(x)
        out = torch.empty(x.shape[0], OB_DIMS, dtype=torch.float32, device=x.device)
        L.check(self.lib.nrt_oneblob_fwd(self.h, L.ptr(x), x.shape[0], L.ptr(out), _stream()))
        return out

    def oneblob_bwd(self, x, dout):
        x, dout = _f32c(x), _f32c(dout)
        dx = torch.empty_like(x)
        L.check(self.lib.nrt_oneblob_bwd(self.h, L.ptr(x), x.shape[0], L.ptr(dout), L.ptr(dx), _stream()))
        return dx

    # ---- decode --------------------------------------------------------------------------------
    def decode_fwd(self, P: FieldTensors, x, with_color=True, want_raw=True, want_sdf_uncert=False, want_geo=False):
        x = _f32c(x)
        n = x.shape[0]
        f = dict(dtype=torch.float32, device=x.device)
        raw = torch.empty(n, 5, **f) if (want_raw and with_color) else None
        su = torch.empty(n, 2, **f) if want_sdf_uncert else None
        geo = torch.empty(n, 15, **f) if want_geo else None
        cp = P.c_params()
        L.check(self.lib.nrt_decode_fwd(self.h, C.byref(cp), L.ptr(x), n, int(with_color), L.ptr(raw), L.ptr(su), L.ptr(geo),
                                        _stream()))
        return raw, su, geo

    def decode_bwd(self, P: FieldTensors, x, draw, G: FieldTensors):
        x, draw = _f32c(x), _f32c(draw)
        n = x.shape[0]
        ws = torch.empty(n * ENC_DIMS, dtype=torch.float32, device=x.device)
        cp, cg = P.c_params(), G.c_grads()
        L.check(self.lib.nrt_decode_bwd(self.h, C.byref(cp), L.ptr(x), n, L.ptr(draw), C.byref(cg), L.ptr(ws), _stream()))

    def map_volumes(self, P: FieldTensors, voxel_size: float):
        """get_map_volumes (src/slam/coslam/coslam_utils.py:58-97) on the device -> (uncert_vol, sdf_vol) device tensors."""
        b = self.bound
        dims = [round(float(b[i, 1] - b[i, 0]) / voxel_size + 0.0005) + 1 for i in range(3)]     # getVoxels, tp/utils.py:36-43
        dev = P.grid.device
        unc = torch.empty(dims, dtype=torch.float32, device=dev)
        sdf = torch.empty(dims, dtype=torch.float32, device=dev)
        cp = P.c_params()
        L.check(self.lib.nrt_map_volumes(self.h, C.byref(cp), (C.c_int32 * 3)(*dims), L.ptr(unc), L.ptr(sdf), _stream()))
        return unc, sdf

    # ---- rays ----------------------------------------------------------------------------------
    def sample_z(self, target_d, u=None, perturb=None, seed=0):
        td = _f32c(target_d).reshape(-1)
        B = td.shape[0]
        z = torch.empty(B, self.S, dtype=torch.float32, device=td.device)
        perturb = self.perturb if perturb is None else int(perturb)
        L.check(self.lib.nrt_sample_z(self.h, L.ptr(td), B, L.ptr(_f32c(u)) if u is not None else None, perturb, seed, L.ptr(z),
                                      _stream()))
        return z

    def composite_fwd(self, raw, z, want_weights=True):
        raw, z = _f32c(raw), _f32c(z)
        B, S = z.shape
        out = RenderBuffers(B, S, z.device, per_sample=False, weights=want_weights)
        cs = out.c_struct()
        L.check(self.lib.nrt_composite_fwd(self.h, L.ptr(raw), L.ptr(z), B, S, C.byref(cs), _stream()))
        return out

    def render_fwd(self, P: FieldTensors, rays_o, rays_d, target_d, out: RenderBuffers, z_in=None, u=None, perturb=None, seed=0):
        rays_o, rays_d = _f32c(rays_o), _f32c(rays_d)
        td = _f32c(target_d).reshape(-1) if target_d is not None else None
        perturb = self.perturb if perturb is None else int(perturb)
        cp, cs = P.c_params(), out.c_struct()
        L.check(self.lib.nrt_render_fwd(self.h, C.byref(cp), L.ptr(rays_o), L.ptr(rays_d), L.ptr(td), rays_o.shape[0],
                                        L.ptr(_f32c(z_in)) if z_in is not None else None,
                                        L.ptr(_f32c(u)) if u is not None else None, perturb, seed, C.byref(cs), _stream()))
        return out

    def render_fwd_stats(self, P: FieldTensors, rays_o, rays_d, target_rgb, target_d, out: RenderBuffers, stats, u=None,
                         perturb=None, seed=0, seed_step=None, losses=None):
        """render_fwd + loss_partial in one launch (training path)."""
        rays_o, rays_d = _f32c(rays_o), _f32c(rays_d)
        perturb = self.perturb if perturb is None else int(perturb)
        cp, cs = P.c_params(), out.c_struct()
        L.check(self.lib.nrt_render_fwd_stats(self.h, C.byref(cp), L.ptr(rays_o), L.ptr(rays_d), L.ptr(_f32c(target_rgb)),
                                              L.ptr(_f32c(target_d).reshape(-1)), rays_o.shape[0],
                                              L.ptr(_f32c(u)) if u is not None else None, perturb, seed,
                                              L.ptr(seed_step) if seed_step is not None else None, C.byref(cs),
                                              L.ptr(stats), L.ptr(losses) if losses is not None else None, _stream()))
        return out

    def new_stats(self, device):
        return torch.zeros(self.lib.nrt_loss_stats_bytes() // 8, dtype=torch.float64, device=device)

    def loss_partial(self, out: RenderBuffers, target_rgb, target_d, stats):
        cs = out.c_struct()
        L.check(self.lib.nrt_loss_partial(self.h, C.byref(cs), L.ptr(_f32c(target_rgb)), L.ptr(_f32c(target_d).reshape(-1)), out.B,
                                          L.ptr(stats), _stream()))

    def loss_finalize(self, stats, losses):
        L.check(self.lib.nrt_loss_finalize(self.h, L.ptr(stats), L.ptr(losses), _stream()))

    def render_bwd(self, P: FieldTensors, rays_o, rays_d, target_rgb, target_d, out: RenderBuffers, stats, loss_grad,
                   G: FieldTensors, workspace=None):
        B = out.B
        need = self.lib.nrt_render_bwd_workspace(self.h, B)
        if workspace is None or workspace.numel() * 4 < need:
            workspace = torch.empty(need // 4, dtype=torch.float32, device=out.rgb.device)
        cp, cg, cs = P.c_params(), G.c_grads(), out.c_struct()
        L.check(self.lib.nrt_render_bwd(self.h, C.byref(cp), L.ptr(_f32c(rays_o)), L.ptr(_f32c(rays_d)), L.ptr(_f32c(target_rgb)),
                                        L.ptr(_f32c(target_d).reshape(-1)), B, C.byref(cs), L.ptr(stats), L.ptr(_f32c(loss_grad)),
                                        C.byref(cg), L.ptr(workspace), _stream()))
        return workspace

    # ---- smoothness / optimiser ------------------------------------------------------------------
    def smooth_workspace(self, n, device):
        return torch.empty(self.lib.nrt_smooth_workspace(self.h, n) // 4, dtype=torch.float32, device=device)

    def smooth_fwd_bwd(self, grid, rand6, n, voxel, margin, loss_scale, loss_out, dgrid, workspace, part=0, n_parts=1):
        L.check(self.lib.nrt_smooth_fwd_bwd(self.h, L.ptr(_f32c(grid)), L.ptr(rand6), int(n), float(voxel), float(margin),
                                            float(loss_scale), L.ptr(loss_out), L.ptr(dgrid), L.ptr(workspace), int(part),
                                            int(n_parts), _stream()))

    def adam_step(self, p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, zero_grad=False, step_dev=None):
        L.check(self.lib.nrt_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), int(step), L.ptr(step_dev), lr, beta1,
                                       beta2, eps, weight_decay, int(zero_grad), _stream()))

    def adam_step_groups(self, theta, grad, exp_avg, exp_avg_sq, groups, zero_grad=True):
        """groups: list of (begin, end, lr, beta1, beta2, eps, weight_decay, step_dev tensor, enabled) over the flat vectors."""
        arr = (L.NrtAdamGroup * len(groups))()
        for a, (b, e, lr, b1, b2, eps, wd, step_dev, en) in zip(arr, groups):
            a.begin, a.end, a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = b, e, lr, b1, b2, eps, wd
            a.step_dev, a.enabled = L.ptr(step_dev), int(en)
        L.check(self.lib.nrt_adam_step_groups(L.ptr(theta), L.ptr(grad), L.ptr(exp_avg), L.ptr(exp_avg_sq), arr, len(groups),
                                              int(zero_grad), _stream()))

    def iteration_begin(self, map_counter, uncert_counter=None, seed=0, rand6=None):
        L.check(self.lib.nrt_iteration_begin(L.ptr(map_counter), L.ptr(uncert_counter) if uncert_counter is not None else None,
                                             int(seed), L.ptr(rand6) if rand6 is not None else None, _stream()))

    def step_begin(self, counter, seed=0, rand6=None, delta=1):
        L.check(self.lib.nrt_step_begin(L.ptr(counter), int(delta), int(seed), L.ptr(rand6) if rand6 is not None else None, _stream()))

    def counter_add(self, counter, delta=1):
        L.check(self.lib.nrt_counter_add(L.ptr(counter), int(delta), _stream()))
