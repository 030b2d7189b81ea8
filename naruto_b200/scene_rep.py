"""Drop-in scene model: same Python surface as the reference's JointEncodingNaruto
(src/slam/coslam/model/scene_rep.py:27-287, base class third_parties/coslam/model/scene_rep.py:10-319), with the
work done by the sm_100a kernels behind libnaruto_b200.so.

A maintainer swaps the import at src/slam/coslam/coslam.py:22
    from naruto_b200.scene_rep import JointEncodingNaruto as JointEncoding
and everything CoSLAMNaruto does with `self.model` keeps working: `.forward`, `.render_rays`, `.query_sdf`,
`.query_color`, `.embed_fn.parameters()`, `.decoder.parameters()`, `.get_uncert_grid()`, `state_dict()` keys.

No torch fallback: every method below launches CUDA kernels through the C-ABI and raises on a CPU tensor.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .field import ENC_DIMS, OB_DIMS, FieldPlan, FieldTensors, RenderBuffers


def _require_cuda(t, what):
    if not t.is_cuda:
        raise L.NrtError(f'naruto_b200: {what} must be a CUDA tensor -- there is no CPU path')


# ------------------------------------------------------------------------------------------------
# lower seam: the two tcnn encodings as nn.Modules (tp/model/encodings.py:31-46, 61-71)
# ------------------------------------------------------------------------------------------------
class _HashEncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, grid, plan):
        ctx.plan = plan
        ctx.save_for_backward(x, grid)
        return plan.encode_fwd(grid, x)

    @staticmethod
    def backward(ctx, dout):
        x, grid = ctx.saved_tensors
        dgrid = torch.zeros_like(grid) if ctx.needs_input_grad[1] else None
        dx = ctx.plan.encode_bwd(grid, x, dout.contiguous(), dgrid, want_dx=ctx.needs_input_grad[0])
        return dx, dgrid, None


class _OneBlobFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan):
        ctx.plan = plan
        ctx.save_for_backward(x)
        return plan.oneblob_fwd(x)

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        return ctx.plan.oneblob_bwd(x, dout.contiguous()), None


class HashGridEncoding(nn.Module):
    """tcnn.Encoding(otype='HashGrid'): one flat fp32 `.params`, `.n_output_dims`, __call__([N,3]) -> [N,32]."""

    def __init__(self, plan: FieldPlan, seed=1337):
        super().__init__()
        self.plan = plan
        self.n_input_dims = 3
        self.n_output_dims = ENC_DIMS
        g = torch.Generator().manual_seed(seed)
        # tcnn initialises grid parameters U(-1e-4, 1e-4)
        self.params = nn.Parameter((torch.rand(plan.n_grid_floats, generator=g) * 2 - 1) * 1e-4)

    def forward(self, x):
        _require_cuda(x, 'encoding input')
        return _HashEncodeFn.apply(x.float().contiguous(), self.params, self.plan)


class OneBlobEncoding(nn.Module):
    """tcnn.Encoding(otype='OneBlob', n_bins=16): no parameters (an empty `.params` like tcnn's)."""

    def __init__(self, plan: FieldPlan):
        super().__init__()
        self.plan = plan
        self.n_input_dims = 3
        self.n_output_dims = OB_DIMS
        self.params = nn.Parameter(torch.zeros(0))

    def forward(self, x):
        _require_cuda(x, 'encoding input')
        return _OneBlobFn.apply(x.float().contiguous(), self.plan)


# ------------------------------------------------------------------------------------------------
# decoder containers: parameters only (same module tree => same state_dict keys as the reference)
# ------------------------------------------------------------------------------------------------
class _Net(nn.Module):
    def __init__(self, n_in, n_hidden, n_out):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(n_in, n_hidden, bias=False), nn.ReLU(inplace=True), nn.Linear(n_hidden, n_out, bias=False))


class ColorSDFNet_v2_Naruto(nn.Module):
    """Parameter container mirroring src/slam/coslam/model/decoder.py:81-116 (color_net registered first)."""

    def __init__(self, config, input_ch=ENC_DIMS, input_ch_pos=OB_DIMS):
        super().__init__()
        d = config['decoder']
        self.color_net = _Net(input_ch_pos + d['geo_feat_dim'], d['hidden_dim_color'], 3)
        self.sdf_net = _Net(input_ch + input_ch_pos, d['hidden_dim'], 1 + d['geo_feat_dim'])


# ------------------------------------------------------------------------------------------------
# autograd bridges
# ------------------------------------------------------------------------------------------------
class _DecodeFn(torch.autograd.Function):
    """query_color_sdf with gradients w.r.t. the six parameter tensors."""

    @staticmethod
    def forward(ctx, x, plan, grid, w1, w2, w3, w4, uncert):
        P = FieldTensors(grid, w1, w2, w3, w4, uncert)
        raw, _, _ = plan.decode_fwd(P, x, with_color=True)
        ctx.plan = plan
        ctx.save_for_backward(x, grid, w1, w2, w3, w4, uncert)
        return raw

    @staticmethod
    def backward(ctx, draw):
        x, *params = ctx.saved_tensors
        P = FieldTensors(*params)
        G = FieldTensors(*[torch.zeros_like(t) for t in params])
        ctx.plan.decode_bwd(P, x, draw.contiguous(), G)
        return (None, None) + tuple(G.as_list())


class _TrainBuffers:
    """Persistent per-batch-size buffers of the training forward (render outputs incl. the saved features / masks, loss
    statistics, backward workspace, one flat gradient buffer): allocated once per batch size and reused, so an iteration
    through the autograd path allocates nothing large (round 1 zero-filled a fresh 67 MB feature buffer and six gradient
    tensors per call).  `version` guards against a second forward overwriting what a pending backward still needs."""

    def __init__(self, plan, B, device, sizes, shapes):
        self.out = RenderBuffers(B, plan.S, device, per_sample=True, feat=True)
        self.stats = plan.new_stats(device)
        self.ws = torch.empty(plan.lib.nrt_render_bwd_workspace(plan.h, B) // 4, dtype=torch.float32, device=device)
        self.sizes, self.shapes = sizes, shapes
        self.version = 0

    def new_grads(self, device):
        flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=device)     # one memset; autograd owns the result
        out, off = [], 0
        for n, shp in zip(self.sizes, self.shapes):
            out.append(flat[off:off + n].view(shp))
            off += n
        return out


class _RenderLossFn(torch.autograd.Function):
    """forward(): ONE launch -- fused render + loss statistics + finalized losses (nrt_render_fwd_stats);
    backward(): nrt_render_bwd (composite_bwd, the TMA-fed MLP-backward / scatter kernel, the weight-gradient reduction)."""

    @staticmethod
    def forward(ctx, plan, bufs, rays_o, rays_d, target_rgb, target_d, u, seed, grid, w1, w2, w3, w4, uncert):
        P = FieldTensors(grid, w1, w2, w3, w4, uncert)
        bufs.version += 1
        losses = torch.empty(L.N_LOSS, dtype=torch.float32, device=rays_o.device)
        plan.render_fwd_stats(P, rays_o, rays_d, target_rgb, target_d, bufs.out, bufs.stats, u=u, seed=seed, losses=losses)
        ctx.plan, ctx.bufs, ctx.version = plan, bufs, bufs.version
        ctx.save_for_backward(rays_o, rays_d, target_rgb, target_d, grid, w1, w2, w3, w4, uncert)
        rgb, depth = bufs.out.rgb.clone(), bufs.out.depth.clone()       # the caller keeps these; the buffers are reused
        ctx.mark_non_differentiable(rgb, depth)
        return losses, rgb, depth

    @staticmethod
    def backward(ctx, dlosses, _drgb, _ddepth):
        rays_o, rays_d, target_rgb, target_d, *params = ctx.saved_tensors
        bufs = ctx.bufs
        if bufs.version != ctx.version:
            raise L.NrtError('naruto_b200: backward() of a training forward whose saved activations were overwritten by a later '
                             'forward() of the same batch size (call backward before the next forward, as the reference\'s '
                             'mapping loop does)')
        P = FieldTensors(*params)
        G = FieldTensors(*bufs.new_grads(rays_o.device))
        ctx.plan.render_bwd(P, rays_o, rays_d, target_rgb, target_d, bufs.out, bufs.stats, dlosses.contiguous(), G, workspace=bufs.ws)
        return (None,) * 8 + tuple(G.as_list())


# ------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------
class JointEncodingNaruto(nn.Module):
    def __init__(self, config, bound_box, uncert_voxel=0.1):
        super().__init__()
        self.config = config
        self.bounding_box = bound_box
        self.plan = FieldPlan(config, bound_box, uncert_voxel=uncert_voxel)
        self.resolution_sdf = self.plan.resolution_sdf
        print('SDF resolution:', self.resolution_sdf)
        # get_encoding (tp/model/scene_rep.py:35-47)
        self.embedpos_fn, self.input_ch_pos = OneBlobEncoding(self.plan), OB_DIMS
        self.embed_fn, self.input_ch = HashGridEncoding(self.plan), ENC_DIMS
        # get_decoder (src/slam/coslam/model/scene_rep.py:38-47): the aliases re-register the sub-nets, which is
        # why the reference's state_dict carries duplicate color_net.* / sdf_net.* keys
        self.decoder = ColorSDFNet_v2_Naruto(config, input_ch=self.input_ch, input_ch_pos=self.input_ch_pos)
        self.color_net = self.decoder.color_net
        self.sdf_net = self.decoder.sdf_net
        self.act_uncertainty = nn.Softplus()
        self._seed = 0
        self._train_bufs = {}                  # (B, S, device) -> _TrainBuffers of the autograd training path
        self.strict_uncert_assert = False      # True => reproduce `assert uncert_map.min() > 0` (costs a host sync)

    # ---- parameters ----------------------------------------------------------------------------
    def get_uncert_grid(self, voxel_size):
        """src/slam/coslam/model/scene_rep.py:49-56."""
        dims = [round((self.bounding_box[i, 1] - self.bounding_box[i, 0]).item() / voxel_size + 0.0005) + 1 for i in range(3)]
        if dims != self.plan.uncert_dims:
            self.plan = FieldPlan(self.config, self.bounding_box, uncert_voxel=voxel_size)
            self.embed_fn.plan = self.embedpos_fn.plan = self.plan
        # the reference hard-codes device="cuda" here (its model always lives there); follow the model instead so that the class
        # also constructs where there is no GPU (host-side tests of the Mapper wiring)
        self.uncert_grid = nn.Parameter(torch.ones(dims, device=self.embed_fn.params.device).float() * 3)
        self.cache_uncert = np.zeros(dims, dtype=np.float32)
        return self.uncert_grid

    def _tensors(self):
        if not hasattr(self, 'uncert_grid'):
            raise L.NrtError('uncert_grid is not initialised: call get_uncert_grid(voxel_size) first '
                             '(CoSLAMNaruto.init_uncert_grid_optim does)')
        return FieldTensors(self.embed_fn.params, self.decoder.sdf_net.model[0].weight, self.decoder.sdf_net.model[2].weight,
                            self.decoder.color_net.model[0].weight, self.decoder.color_net.model[2].weight, self.uncert_grid)

    def _next_seed(self):
        self._seed += 1
        return self._seed

    # ---- point queries ---------------------------------------------------------------------------
    def calc_embedding(self, inputs):
        """[uncert | hash features] (src/slam/coslam/model/scene_rep.py:58-64), no gradients."""
        _require_cuda(inputs, 'inputs')
        P = self._tensors()
        embed = self.plan.encode_fwd(P.grid, inputs)
        _, su, _ = self.plan.decode_fwd(P, inputs, with_color=False, want_raw=False, want_sdf_uncert=True)
        return torch.cat([su[:, 1:2], embed], dim=1)

    def query_sdf(self, query_points, return_geo=False, embed=False, return_uncert=False):
        """src/slam/coslam/model/scene_rep.py:98-130 (query_points already normalised to the bound)."""
        _require_cuda(query_points, 'query_points')
        flat = torch.reshape(query_points, [-1, query_points.shape[-1]])
        lead = list(query_points.shape[:-1])
        if embed:
            return torch.reshape(self.embed_fn(flat), lead + [ENC_DIMS])
        _, su, geo = self.plan.decode_fwd(self._tensors(), flat, with_color=False, want_raw=False, want_sdf_uncert=True,
                                          want_geo=return_geo)
        sdf = torch.reshape(su, lead + [2]) if return_uncert else torch.reshape(su[:, 0], lead)
        if not return_geo:
            return sdf
        return sdf, torch.reshape(geo, lead + [15])

    def query_color_sdf(self, query_points):
        """src/slam/coslam/model/scene_rep.py:132-148 -> raw [N,5]; differentiable w.r.t. the parameters."""
        _require_cuda(query_points, 'query_points')
        flat = torch.reshape(query_points, [-1, query_points.shape[-1]]).float().contiguous()
        return _DecodeFn.apply(flat, self.plan, *self._tensors().as_list())

    def query_color(self, query_points):
        return torch.sigmoid(self.query_color_sdf(query_points)[..., :3])

    def run_network(self, inputs):
        """tp/model/scene_rep.py:160-178."""
        flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
        if self.config['grid']['tcnn_encoding']:
            bb = self.bounding_box.to(flat)
            flat = (flat - bb[:, 0]) / (bb[:, 1] - bb[:, 0])
        out = self.query_color_sdf(flat)
        return torch.reshape(out, list(inputs.shape[:-1]) + [out.shape[-1]])

    # ---- compositing -----------------------------------------------------------------------------
    def sdf2weights(self, sdf, z_vals, args=None):
        """tp/model/scene_rep.py:64-84 (no gradients through this entry point)."""
        raw = torch.zeros(*sdf.shape, 5, device=sdf.device)
        raw[..., 3] = sdf
        return self.plan.composite_fwd(raw, z_vals).weights

    def raw2outputs(self, raw, z_vals, white_bkgd=False):
        """src/slam/coslam/model/scene_rep.py:66-96 (no gradients through this entry point)."""
        if white_bkgd:
            raise L.NrtError('white_bkgd=True is not implemented')
        o = self.plan.composite_fwd(raw, z_vals)
        return o.rgb, o.disp, o.acc, o.weights, o.depth, o.depth_var, o.uncert

    # ---- inactive-at-shipped-config surface: present, explicit ---------------------------------
    def get_resolution(self):
        """tp/model/scene_rep.py:20-35: resolution_sdf from grid.voxel_sdf (done by FieldPlan at construction)."""
        self.resolution_sdf = self.plan.resolution_sdf
        return self.resolution_sdf

    def render_surface_color(self, rays_o, normal):
        """tp/model/scene_rep.py:180-196: colour of surface points from n_range_d samples along +-trunc of the normal.
        (In the reference this method unpacks six values from JointEncodingNaruto.raw2outputs, which returns seven, so it
        raises ValueError there; `training.render_color` is False at every shipped config.  Served here by the same
        kernels as render_rays: point decode + compositor.)"""
        _require_cuda(rays_o, 'rays_o')
        with torch.no_grad():
            t = self.config['training']
            z = torch.linspace(-t['trunc'], t['trunc'], steps=t['n_range_d']).to(rays_o).repeat(rays_o.shape[0], 1)
            pts = rays_o[..., :] + normal[..., None, :] * z[..., :, None]
            raw = self.run_network(pts)
            return self.plan.composite_fwd(raw, z, want_weights=False).rgb

    @staticmethod
    def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5):
        """tp/model/utils.py:29-68 (hierarchical importance sampling).  Only reached with training.n_importance > 0, which
        FieldPlan rejects at construction (0 at every shipped config): there is no kernel for it and no torch fallback."""
        raise L.NrtError('sample_pdf / n_importance > 0 is not implemented by naruto_b200 (n_importance is 0 at every shipped '
                         'NARUTO config); FieldPlan rejects such a config at construction')

    # ---- rays --------------------------------------------------------------------------------------
    def render_rays(self, rays_o, rays_d, target_d=None, u=None, z_vals=None):
        """src/slam/coslam/model/scene_rep.py:150-225.  `u` ([B,S] uniform draws) / `z_vals` are optional parity hooks:
        by default the kernel draws its own Philox stream where the reference calls torch.rand."""
        _require_cuda(rays_o, 'rays_o')
        if target_d is None and z_vals is None:
            # the reference needs training.n_samples here, which the shipped configs comment out (KeyError)
            raise KeyError('n_samples')
        with torch.no_grad():
            B = rays_o.shape[0]
            out = RenderBuffers(B, self.plan.S, rays_o.device, per_sample=True)
            self.plan.render_fwd(self._tensors(), rays_o, rays_d, target_d, out, z_in=z_vals, u=u, seed=self._next_seed())
        return {'rgb': out.rgb, 'depth': out.depth, 'disp_map': out.disp, 'acc_map': out.acc, 'depth_var': out.depth_var,
                'z_vals': out.z_vals, 'raw': out.raw, 'uncert_map': out.uncert}

    def forward(self, rays_o, rays_d, target_rgb, target_d, global_step=0, u=None):
        """src/slam/coslam/model/scene_rep.py:227-287."""
        if not self.training:
            return self.render_rays(rays_o, rays_d, target_d=target_d, u=u)
        _require_cuda(rays_o, 'rays_o')
        if rays_o.requires_grad or rays_d.requires_grad:
            # the reference lets dL/d rays flow into the pose parameters (src/slam/coslam/coslam.py:265,342-344); this
            # library has no d rays_o / d rays_d output, and NARUTO runs with tracking / pose refinement disabled
            # (poses are ground truth, `cur_rot`/`cur_trans` never receive an optimiser step): refuse instead of
            # silently detaching.
            raise L.NrtError('naruto_b200: rays_o / rays_d require grad, but the fused renderer has no pose gradient '
                             '(NARUTO maps with fixed poses); pass detached rays')
        f = lambda t: t.detach().float().contiguous()
        P = self._tensors()
        B = rays_o.shape[0]
        key = (B, self.plan.S, rays_o.device)
        if key not in self._train_bufs:
            if len(self._train_bufs) >= 4:
                self._train_bufs.pop(next(iter(self._train_bufs)))
            tl = P.as_list()
            self._train_bufs[key] = _TrainBuffers(self.plan, B, rays_o.device, [t.numel() for t in tl], [tuple(t.shape) for t in tl])
        losses, rgb, depth = _RenderLossFn.apply(self.plan, self._train_bufs[key], f(rays_o), f(rays_d), f(target_rgb),
                                                 f(target_d), u, self._next_seed(), *P.as_list())
        if self.strict_uncert_assert:
            assert losses[L.LOSS_UNCERT_MIN].item() > 0
        return {'rgb': rgb, 'depth': depth, 'rgb_loss': losses[L.LOSS_RGB], 'depth_loss': losses[L.LOSS_DEPTH],
                'sdf_loss': losses[L.LOSS_SDF], 'fs_loss': losses[L.LOSS_FS], 'psnr': losses[L.LOSS_PSNR:L.LOSS_PSNR + 1].detach(),
                'uncert_loss': losses[L.LOSS_UNCERT]}
