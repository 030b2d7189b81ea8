"""Fused mapping iteration: the body of CoSLAMNaruto.global_BA's loop (src/slam/coslam/coslam.py:364-399) --
model.forward, get_loss_from_ret(smooth=True), loss.backward(), map_optimizer.step()/zero_grad() and the every-5th
uncert_optim.step()/zero_grad() -- as a fixed sequence of launches into libnaruto_b200.so, captured once into a
CUDA graph and replayed.

Data parallel: each rank renders its own shard of the ray batch; the loss statistics (11 doubles) are
all-reduced before the backward pass because every loss is a ratio of global sums, and the flat gradient buffer
[grid | w1 | w2 | w3 | w4 | uncert] is all-reduced (NCCL over NVLink) before an identical Adam step on every rank.
"""
import torch

from . import _lib as L
from .field import FieldPlan, FieldTensors, RenderBuffers
from .parallel import reduce_grads, reduce_stats


class MappingStep:
    def __init__(self, plan: FieldPlan, cfg: dict, n_rays: int, device, init: FieldTensors = None, process_group=None,
                 use_graph: bool = True, smooth: bool = True):
        self.plan, self.cfg, self.B, self.dev = plan, cfg, int(n_rays), torch.device(device)
        self.pg = process_group
        self.rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        t, mp = cfg['training'], cfg['mapping']
        if int(mp.get('map_accum_step', 1)) != 1 or int(mp.get('map_wait_step', 0)) != 0:
            # src/slam/coslam/coslam.py:368-376: the reference steps the optimiser every map_accum_step iterations after
            # map_wait_step; the fused iteration steps every time (1 / 0 at every shipped config)
            raise L.NrtError('MappingStep implements mapping.map_accum_step == 1 and map_wait_step == 0 (the shipped values)')
        self.lr_decoder, self.lr_embed = float(mp['lr_decoder']), float(mp['lr_embed'])
        self.smooth_on = bool(smooth) and t['smooth_weight'] > 0
        self.smooth_w, self.smooth_n = float(t['smooth_weight']), int(t['smooth_pts'])
        self.smooth_vox, self.smooth_margin = float(t['smooth_vox']), float(t['smooth_margin'])
        f32 = dict(dtype=torch.float32, device=self.dev)
        ud = plan.uncert_dims
        sizes = [plan.n_grid_floats, 32 * 80, 16 * 32, 32 * 63, 3 * 32, ud[0] * ud[1] * ud[2]]
        shapes = [(plan.n_grid_floats,), (32, 80), (16, 32), (32, 63), (3, 32), tuple(ud)]
        self.n_grid, self.n_dec, self.n_unc = sizes[0], sum(sizes[1:5]), sizes[5]
        total = sum(sizes)
        self.theta = torch.zeros(total, **f32)          # parameters, one flat buffer
        # gradients, same layout, plus one trailing slot for this rank's part of the smoothness loss: the whole buffer is the
        # all-reduce bucket, so the loss value is summed across ranks for free
        self.bucket = torch.zeros(total + 1, **f32)
        self.grad = self.bucket[:total]
        self.exp_avg = torch.zeros(total, **f32)
        self.exp_avg_sq = torch.zeros(total, **f32)

        def views(buf):
            out, off = [], 0
            for n, shp in zip(sizes, shapes):
                out.append(buf[off:off + n].view(shp))
                off += n
            return FieldTensors(*out)

        self.P, self.G = views(self.theta), views(self.grad)
        if init is not None:
            with torch.no_grad():
                for dst, src in zip(self.P.as_list(), init.as_list()):
                    dst.copy_(src.detach().to(self.dev))
        # static buffers (graph inputs / outputs)
        B = self.B
        self.inbuf = torch.zeros(10 * B, **f32)         # packed [o | d | rgb | depth]: one H2D copy per iteration
        self.rays_o = self.inbuf[0:3 * B].view(B, 3)
        self.rays_d = self.inbuf[3 * B:6 * B].view(B, 3)
        self.target_rgb = self.inbuf[6 * B:9 * B].view(B, 3)
        self.target_d = self.inbuf[9 * B:10 * B].view(B, 1)
        self.out = RenderBuffers(self.B, plan.S, self.dev, per_sample=True, feat=True)
        self.u = torch.zeros(self.B, plan.S, **f32)
        self.rand6 = torch.zeros(6, **f32)
        self.stats = plan.new_stats(self.dev)
        self.losses = torch.zeros(L.N_LOSS, **f32)
        self.smooth_loss = self.bucket[total:total + 1]
        self.loss_grad = torch.tensor([t['rgb_weight'], t['depth_weight'], t['sdf_weight'], t['fs_weight'],
                                       t.get('uncert_weight', 0.0)], **f32)
        self.ws_bwd = torch.empty(plan.lib.nrt_render_bwd_workspace(plan.h, self.B) // 4, **f32)
        self.ws_smooth = plan.smooth_workspace(self.smooth_n, self.dev)
        self.map_step = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.unc_step = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.it = 0
        self.use_graph = use_graph
        self.external_random = False     # test hook: keep caller-written self.u / self.rand6 instead of drawing
        # two Philox keys: the stratified jitter is per ray, so its stream is rank-offset; the smoothness lattice
        # draw (rand6) must be THE SAME on every rank -- each rank takes one slab of one common lattice, and only then do the
        # slabs add up to the reference's smoothness() term in the gradient all-reduce
        self.base_seed = 0x9E3779B9
        self.seed = self.base_seed + 7919 * self.rank
        self._graphs = {}
        self.launches_per_iter = {False: 0, True: 0}

    # -------------------------------------------------------------------------------------------
    def _body(self, with_uncert_step: bool):
        """The launches of one iteration on the current stream.  Returns how many kernels of ours it launched."""
        p, n = self.plan, 0
        fused_losses = self.losses if self.world == 1 else None      # one shard: the render kernel's last CTA finalizes the losses
        if self.external_random:                            # test hook: caller-written self.u / self.rand6 (the reference's draws)
            p.counter_add(self.map_step, 1); n += 1
            p.render_fwd_stats(self.P, self.rays_o, self.rays_d, self.target_rgb, self.target_d, self.out, self.stats, u=self.u,
                               losses=fused_losses); n += 1
        else:
            # the reference's torch.rand(z_vals.shape), torch.rand(3), torch.rand((1,1,1,3)) draws, made on the device from Philox
            # keyed by (seed, step counter): nothing host-side changes between graph replays
            p.step_begin(self.map_step, self.base_seed, self.rand6 if self.smooth_on else None); n += 1
            p.render_fwd_stats(self.P, self.rays_o, self.rays_d, self.target_rgb, self.target_d, self.out, self.stats, u=None,
                               seed=self.seed, seed_step=self.map_step, losses=fused_losses); n += 1
        if fused_losses is None:                            # N > 1: the losses are ratios of GLOBAL sums
            reduce_stats(self.stats, self.pg)
            p.loss_finalize(self.stats, self.losses); n += 1
        p.render_bwd(self.P, self.rays_o, self.rays_d, self.target_rgb, self.target_d, self.out, self.stats, self.loss_grad,
                     self.G, workspace=self.ws_bwd); n += 2
        if self.smooth_on:                                  # ray-independent term: every rank takes one slab of the lattice
            p.smooth_fwd_bwd(self.P.grid, self.rand6, self.smooth_n, self.smooth_vox, self.smooth_margin, self.smooth_w,
                             self.smooth_loss, self.G.grid, self.ws_smooth, part=self.rank, n_parts=self.world); n += 2
        reduce_grads(self.bucket, self.pg)
        ng, nd = self.n_grid, self.n_dec
        # create_optimizer (src/slam/coslam/coslam.py:409-419): decoder group wd=1e-6, grid group eps=1e-15, betas (0.9,0.99)
        p.adam_step(self.theta[:ng], self.grad[:ng], self.exp_avg[:ng], self.exp_avg_sq[:ng], 0, self.lr_embed, 0.9, 0.99,
                    1e-15, 0.0, zero_grad=True, step_dev=self.map_step); n += 1
        p.adam_step(self.theta[ng:ng + nd], self.grad[ng:ng + nd], self.exp_avg[ng:ng + nd], self.exp_avg_sq[ng:ng + nd], 0,
                    self.lr_decoder, 0.9, 0.99, 1e-8, 1e-6, zero_grad=True, step_dev=self.map_step); n += 1
        if with_uncert_step:
            # init_uncert_grid_optim (:240-243): Adam(lr=1), stepped and zeroed every 5th iteration (:397-399);
            # in between the uncertainty-grid gradient keeps accumulating
            o = ng + nd
            p.counter_add(self.unc_step, 1); n += 1
            p.adam_step(self.theta[o:], self.grad[o:], self.exp_avg[o:], self.exp_avg_sq[o:], 0, 1.0, 0.9, 0.999, 1e-8, 0.0,
                        zero_grad=True, step_dev=self.unc_step); n += 1
        return n

    def _graph(self, with_uncert_step):
        if with_uncert_step not in self._graphs:
            # warm-up outside capture (lazy attribute/module init inside the library and torch RNG)
            side = torch.cuda.Stream(device=self.dev)
            saved = [b.clone() for b in (self.theta, self.grad, self.exp_avg, self.exp_avg_sq, self.map_step, self.unc_step)]
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._body(with_uncert_step)
            torch.cuda.current_stream().wait_stream(side)
            for b, s in zip((self.theta, self.grad, self.exp_avg, self.exp_avg_sq, self.map_step, self.unc_step), saved):
                b.copy_(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.launches_per_iter[with_uncert_step] = self._body(with_uncert_step)
            for b, s in zip((self.theta, self.grad, self.exp_avg, self.exp_avg_sq, self.map_step, self.unc_step), saved):
                b.copy_(s)
            self._graphs[with_uncert_step] = g
        return self._graphs[with_uncert_step]

    # -------------------------------------------------------------------------------------------
    def release_graphs(self):
        """Destroy the captured CUDA graphs (they hold references to the NCCL communicator: do this before
        torch.distributed.destroy_process_group())."""
        self._graphs.clear()

    def load_rays(self, rays_o, rays_d, target_rgb, target_d):
        """Host (pinned) or device tensors -> the static input buffers (async copies on the current stream)."""
        self.rays_o.copy_(rays_o.reshape(self.B, 3), non_blocking=True)
        self.rays_d.copy_(rays_d.reshape(self.B, 3), non_blocking=True)
        self.target_rgb.copy_(target_rgb.reshape(self.B, 3), non_blocking=True)
        self.target_d.copy_(target_d.reshape(self.B, 1), non_blocking=True)

    def load_packed(self, buf):
        """One packed [10*B] host (pinned) or device buffer, see SyntheticFrame.sample_packed."""
        self.inbuf.copy_(buf, non_blocking=True)

    def step(self, rays_o=None, rays_d=None, target_rgb=None, target_d=None):
        """One mapping iteration.  Returns the device tensor of the five losses (no host sync)."""
        if rays_o is not None:
            self.load_rays(rays_o, rays_d, target_rgb, target_d)
        with_unc = (self.it + 1) % 5 == 0
        if self.use_graph:
            self._graph(with_unc).replay()
        else:
            self.launches_per_iter[with_unc] = self._body(with_unc)
        self.it += 1
        return self.losses

    def total_loss(self):
        """get_loss_from_ret's scalar (host sync)."""
        l = self.losses[:5].double().cpu()
        w = self.loss_grad.double().cpu()
        tot = float((l * w).sum())
        if self.smooth_on:
            tot += self.smooth_w * float(self.smooth_loss.item())
        return tot
