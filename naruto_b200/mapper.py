"""Fused mapping iteration: the body of CoSLAMNaruto.global_BA's loop (src/slam/coslam/coslam.py:364-399) --
model.forward, get_loss_from_ret(smooth=True), loss.backward(), map_optimizer.step()/zero_grad() and the every-5th
uncert_optim.step()/zero_grad() -- as a fixed sequence of 8 launches into libnaruto_b200.so (the smoothness pair on a forked
branch), captured once into a CUDA graph and replayed.

Data parallel: each rank renders its own shard of the ray batch; the loss statistics (11 doubles) are
exchanged before the backward pass because every loss is a ratio of global sums, and the flat gradient buffer
[grid | w1 | w2 | w3 | w4 | uncert] is reduce-scattered, stepped and all-gathered by ONE kernel over NVLink peer memory
(csrc/peer.cu; NCCL all-reduces + the same Adam launch where symmetric memory is unavailable or NRT_DP_IMPL=nccl).
"""
import os

import torch

from . import _lib as L
from .field import FieldPlan, FieldTensors, RenderBuffers
from .parallel import PeerExchange, reduce_grads, reduce_stats


class FusedState:
    """Parameters, gradients and Adam moments of the scene model as flat device buffers [grid | w1 | w2 | w3 | w4 | uncert]
    (the gradient buffer is the all-reduce bucket), plus the device-side step counters.  One FusedState can serve several
    MappingStep objects (one per batch size); `bind_model` makes a JointEncodingNaruto's nn.Parameters and the state of its
    torch.optim.Adam optimisers VIEWS of these buffers, so the fused iteration, the autograd path, `query_sdf`,
    `get_map_volumes`, `save_mesh` and `save_ckpt` / `state_dict` all see one and the same set of numbers."""

    def __init__(self, plan: FieldPlan, device, init: FieldTensors = None, process_group=None):
        self.plan, self.dev = plan, torch.device(device)
        f32 = dict(dtype=torch.float32, device=self.dev)
        ud = plan.uncert_dims
        self.sizes = [plan.n_grid_floats, 32 * 80, 16 * 32, 32 * 63, 3 * 32, ud[0] * ud[1] * ud[2]]
        self.shapes = [(plan.n_grid_floats,), (32, 80), (16, 32), (32, 63), (3, 32), tuple(ud)]
        self.n_grid, self.n_dec, self.n_unc = self.sizes[0], sum(self.sizes[1:5]), self.sizes[5]
        total = self.total = sum(self.sizes)
        # Data parallel: parameters and the gradient bucket live in NVLink-mapped symmetric memory when it is available, and the
        # two exchanges of the iteration run inside our own kernels (csrc/peer.cu); NRT_DP_IMPL=nccl forces the NCCL all-reduces
        self.peers = None
        self.pg = process_group
        if process_group is not None and torch.distributed.get_world_size(process_group) > 1 \
                and os.environ.get('NRT_DP_IMPL', 'peer') != 'nccl' and self.dev.type == 'cuda':
            try:
                self.peers = PeerExchange(total, self.dev, process_group)
            except Exception as e:       # no P2P / symmetric memory on this system: NCCL path
                import warnings
                warnings.warn(f'naruto_b200: peer-memory exchange unavailable ({e!r}); using NCCL all-reduces')
                self.peers = None
        self.smooth_slot = total
        pad = (total + 3) // 4 * 4
        if self.peers is not None:
            self.theta, self.bucket, self.smooth_slot = self.peers.theta[:total], self.peers.bucket, self.peers.smooth_slot
        else:
            self.theta = torch.zeros(total, **f32)          # parameters, one flat buffer
            # gradients, same layout, plus one trailing slot for this rank's part of the smoothness loss: the whole buffer is the
            # all-reduce bucket, so the loss value is summed across ranks for free
            self.bucket = torch.zeros(total + 1, **f32)
        self.grad = self.bucket[:total]
        self.exp_avg = torch.zeros(pad, **f32)[:total]          # (storage padded to whole float4s for the peer-memory Adam)
        self.exp_avg_sq = torch.zeros(pad, **f32)[:total]
        self.P, self.G = self.views(self.theta), self.views(self.grad)
        self.M, self.V = self.views(self.exp_avg), self.views(self.exp_avg_sq)
        self.map_step = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.unc_step = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.n_map_steps = 0        # host mirrors of the two device counters (no synchronisation needed to know them)
        self.n_unc_steps = 0
        self._bound = None
        if init is not None:
            with torch.no_grad():
                for dst, src in zip(self.P.as_list(), init.as_list()):
                    dst.copy_(src.detach().to(self.dev))

    def views(self, buf):
        out, off = [], 0
        for n, shp in zip(self.sizes, self.shapes):
            out.append(buf[off:off + n].view(shp))
            off += n
        return FieldTensors(*out)

    def persistent(self):
        return (self.theta, self.grad, self.exp_avg, self.exp_avg_sq, self.map_step, self.unc_step)

    # ---- one parameter set for every path ---------------------------------------------------------
    def bind_model(self, model, map_optimizer=None, uncert_optim=None):
        """Alias `model`'s parameters (embed_fn.params, the four decoder weights, uncert_grid) onto this state: their
        current values are copied in once, then `param.data` becomes a view of `theta` and `param.grad` a view of the
        gradient buffer.  If the torch optimisers created by CoSLAMNaruto.create_optimizer / init_uncert_grid_optim
        (src/slam/coslam/coslam.py:409-419, 240-243) are passed, their Adam state (`exp_avg`, `exp_avg_sq`, `step`) is made to
        alias / mirror the fused moments too: `optimizer.state_dict()` (save_ckpt) and `optimizer.step()` (the autograd path
        of first_frame_mapping) stay consistent with fused iterations."""
        params = model._tensors().as_list()
        with torch.no_grad():
            for dst, g, p in zip(self.P.as_list(), self.G.as_list(), params):
                if p.data_ptr() != dst.data_ptr():
                    dst.copy_(p.detach().to(self.dev))
                    p.data = dst
                p.grad = g
        self._bound = (model, map_optimizer, uncert_optim)
        self.sync_optimizers()
        return self

    def owned_ranges(self, rank=None, world=None):
        """Float ranges [lo, hi) of the flat vector whose Adam moments THIS rank maintains under the peer-memory exchange
        (csrc/peer.cu: rank r owns float4s [b4 + n4*r/W, b4 + n4*(r+1)/W) of every parameter group)."""
        if self.peers is None:
            return [(0, self.total)]
        W = self.peers.world if world is None else world
        R = self.peers.rank if rank is None else rank
        out, ng, nd = [], self.n_grid, self.n_dec
        for b, e in ((0, ng), (ng, ng + nd), (ng + nd, self.total)):
            b4, n4 = b // 4, (e + 3) // 4 - b // 4
            out.append((4 * (b4 + n4 * R // W), min(4 * (b4 + n4 * (R + 1) // W), self.total)))
        return out

    def gather_moments(self):
        """Peer-memory data parallelism shards the Adam moments (every rank steps 1/N of each group): make exp_avg / exp_avg_sq
        complete on every rank again, e.g. before `save_ckpt` reads `optimizer.state_dict()`.  One all-reduce of two masked copies."""
        if self.peers is None:
            return
        for m in (self.exp_avg, self.exp_avg_sq):
            full = torch.zeros_like(m)
            for lo, hi in self.owned_ranges():
                full[lo:hi] = m[lo:hi]
            torch.distributed.all_reduce(full, group=self.pg)
            m.copy_(full)

    def sync_optimizers(self, gather=True):
        """Write the fused Adam state into the bound torch optimisers (views for the moments, host step counts); at N > 1 with
        sharded moments, all-gather them first (gather=False skips that)."""
        if gather:
            self.gather_moments()
        if self.peers is not None:
            # the peer-memory iteration clears a rank's [grid | decoder] gradients at the START of the next iteration (no
            # zero-stores over NVLink): leave them as optimizer.zero_grad() would for whoever runs next
            self.grad[:self.n_grid + self.n_dec].zero_()
        if self._bound is None:
            return
        model, map_opt, unc_opt = self._bound
        params = model._tensors().as_list()
        for opt, idxs, nstep in ((map_opt, range(0, 5), self.n_map_steps), (unc_opt, (5,), self.n_unc_steps)):
            if opt is None:
                continue
            owned = {id(p) for grp in opt.param_groups for p in grp['params']}
            for i in idxs:
                p = params[i]
                if id(p) not in owned:
                    continue
                st = opt.state[p]
                st['exp_avg'], st['exp_avg_sq'] = self.M.as_list()[i], self.V.as_list()[i]
                st['step'] = torch.tensor(float(nstep))

    def adopt_optimizer_steps(self):
        """If the bound torch optimisers were stepped outside the fused path (autograd path), take over their step counts."""
        if self._bound is None:
            return
        model, map_opt, unc_opt = self._bound
        params = model._tensors().as_list()

        def count(opt, i):
            if opt is None or params[i] not in opt.state or 'step' not in opt.state[params[i]]:
                return None
            return int(float(opt.state[params[i]]['step']))

        n = count(map_opt, 0)
        if n is not None and n != self.n_map_steps:
            self.n_map_steps = n
            self.map_step.fill_(n)
        n = count(unc_opt, 5)
        if n is not None and n != self.n_unc_steps:
            self.n_unc_steps = n
            self.unc_step.fill_(n)


class MappingStep:
    def __init__(self, plan: FieldPlan, cfg: dict, n_rays: int, device, init: FieldTensors = None, process_group=None,
                 use_graph: bool = True, smooth: bool = True, state: FusedState = None):
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
        self.state = st = state if state is not None else FusedState(plan, device, init=init, process_group=process_group)
        self.peers = st.peers
        self.n_grid, self.n_dec, self.n_unc = st.n_grid, st.n_dec, st.n_unc
        total = st.total
        self.theta, self.bucket, self.grad, self.exp_avg, self.exp_avg_sq = st.theta, st.bucket, st.grad, st.exp_avg, st.exp_avg_sq
        self.P, self.G = st.P, st.G
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
        self.host_in = self.host_losses = None       # pinned host buffers of step_host(), allocated on first use
        self.smooth_loss = self.bucket[st.smooth_slot:st.smooth_slot + 1]
        # N > 1 with peer memory: the sum of the ranks' slab losses lands here (the bucket slot itself is rank-local then)
        self.smooth_total = torch.zeros(1, **f32)
        self.loss_grad = torch.tensor([t['rgb_weight'], t['depth_weight'], t['sdf_weight'], t['fs_weight'],
                                       t.get('uncert_weight', 0.0)], **f32)
        self.ws_bwd = torch.empty(plan.lib.nrt_render_bwd_workspace(plan.h, self.B) // 4, **f32)
        self.ws_smooth = plan.smooth_workspace(self.smooth_n, self.dev)
        self.map_step, self.unc_step = st.map_step, st.unc_step
        self.it = 0
        self.use_graph = use_graph
        self.external_random = False     # test hook: keep caller-written self.u / self.rand6 instead of drawing
        # two Philox keys: the stratified jitter is per ray, so its stream is rank-offset; the smoothness lattice
        # draw (rand6) must be THE SAME on every rank -- each rank takes one slab of one common lattice, and only then do the
        # slabs add up to the reference's smoothness() term in the gradient all-reduce
        self.base_seed = 0x9E3779B9
        self.seed = self.base_seed + 7919 * self.rank
        self._graphs = {}
        # The ray-independent smoothness term runs on a forked branch of the iteration, joined before the optimiser step
        # (NRT_SMOOTH_FORK: 0 in line, 1 forked after the iteration's first launch, 2 forked after the render forward = default), held
        # back by NRT_SMOOTH_STAGGER one-block launches so that the ray path's next kernel (composite_bwd) is resident first
        # and the lattice kernels fill in around it.  Measured on B200 (profiles/r02d_smooth_branch.log), us per iteration at
        # 4096 rays x 128 samples / 2048 x 43: in line 449 / 213; fork 1 426 / 203; fork 2 433 / 187; fork 2 + stagger 2
        # 424 / 186; a branch that starts together with decode_bwd_q (stagger 8) costs more than it hides (453 / 207).
        self.smooth_fork = int(os.environ.get('NRT_SMOOTH_FORK', '2'))
        self.smooth_stagger = int(os.environ.get('NRT_SMOOTH_STAGGER', '2'))
        self._branch = None
        self._zero_branch = None
        self.local_zero = os.environ.get('NRT_DP_LOCAL_ZERO', '1') == '1'
        # profiling aid (tools/probe_dp.py): NRT_STEP_STAMPS=1 puts a one-thread globaltimer stamp after every stage of the
        # iteration, also inside the captured graph, where events cannot time; stamps[k] <-> stamp_names[k]
        self.stamps = torch.zeros(16, dtype=torch.int64, device=self.dev) if os.environ.get('NRT_STEP_STAMPS') == '1' else None
        self.stamp_names, self._stamp_i = [], 0
        self.launches_per_iter = {False: 0, True: 0}

    # -------------------------------------------------------------------------------------------
    def _body(self, with_uncert_step: bool, smooth: bool = None):
        """The launches of one iteration on the current stream.  Returns how many kernels of ours it launched."""
        p, n = self.plan, 0
        smooth = self.smooth_on if smooth is None else (bool(smooth) and self.smooth_w > 0)
        fused_losses = self.losses if self.world == 1 else None      # one shard: the render kernel's last CTA finalizes the losses
        forked = zeroing = False
        self._stamp_i = 0
        self._mark('start')
        if self.external_random:                            # test hook: caller-written self.u / self.rand6 (the reference's draws)
            p.counter_add(self.map_step, 1); n += 1
            p.render_fwd_stats(self.P, self.rays_o, self.rays_d, self.target_rgb, self.target_d, self.out, self.stats, u=self.u,
                               losses=fused_losses); n += 1
        else:
            # the reference's torch.rand(z_vals.shape), torch.rand(3), torch.rand((1,1,1,3)) draws, made on the device from Philox
            # keyed by (seed, step counter): nothing host-side changes between graph replays
            p.iteration_begin(self.map_step, self.unc_step if with_uncert_step else None, self.base_seed,
                              self.rand6 if smooth else None); n += 1
            if self.peers is not None and self.local_zero:
                # every rank clears its own [grid | decoder] gradients, on a branch beside the forward (the peers read them in the
                # previous iteration's optimiser launch, which ended with a barrier): no zero-stores over NVLink
                if self._zero_branch is None:
                    self._zero_branch = torch.cuda.Stream(device=self.dev)
                self._zero_branch.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._zero_branch):
                    self.grad[:self.n_grid + self.n_dec].zero_()
                zeroing = True
            if smooth and self.smooth_fork == 1 and not zeroing:
                n += self._smooth_branch(); forked = True
            p.render_fwd_stats(self.P, self.rays_o, self.rays_d, self.target_rgb, self.target_d, self.out, self.stats, u=None,
                               seed=self.seed, seed_step=self.map_step, losses=fused_losses); n += 1
        if zeroing:
            torch.cuda.current_stream().wait_stream(self._zero_branch)
        self._mark('forward')
        if fused_losses is None:                            # N > 1: the losses are ratios of GLOBAL sums
            if self.peers is not None:                      # exchange + finalize in one launch over peer memory
                self.peers.stats_exchange(self.stats, self.losses); n += 1
            else:
                reduce_stats(self.stats, self.pg)
                p.loss_finalize(self.stats, self.losses); n += 1
        self._mark('stats_exchange')
        if smooth and self.smooth_fork == 2:
            n += self._smooth_branch(); forked = True
        p.render_bwd(self.P, self.rays_o, self.rays_d, self.target_rgb, self.target_d, self.out, self.stats, self.loss_grad,
                     self.G, workspace=self.ws_bwd); n += 3     # composite_bwd, decode_bwd_q, wgrad_reduce
        self._mark('backward')
        if smooth and not forked:                           # ray-independent term: every rank takes one slab of the lattice
            p.smooth_fwd_bwd(self.P.grid, self.rand6, self.smooth_n, self.smooth_vox, self.smooth_margin, self.smooth_w,
                             self.smooth_loss, self.G.grid, self.ws_smooth, part=self.rank, n_parts=self.world); n += 2
        elif forked:
            torch.cuda.current_stream().wait_stream(self._branch)      # join: the optimiser needs both gradients
        self._mark('smooth/join')
        ng, nd = self.n_grid, self.n_dec
        if with_uncert_step and self.external_random:       # (otherwise iteration_begin advanced the uncertainty step counter)
            p.counter_add(self.unc_step, 1); n += 1
        # create_optimizer (src/slam/coslam/coslam.py:409-419): grid group eps=1e-15, decoder group wd=1e-6, betas (0.9,0.99);
        # init_uncert_grid_optim (:240-243): Adam(lr=1), stepped and zeroed every 5th iteration (:397-399) -- in between the
        # uncertainty-grid gradient keeps accumulating
        groups = [(0, ng, self.lr_embed, 0.9, 0.99, 1e-15, 0.0, self.map_step, True),
                  (ng, ng + nd, self.lr_decoder, 0.9, 0.99, 1e-8, 1e-6, self.map_step, True),
                  (ng + nd, ng + nd + self.n_unc, 1.0, 0.9, 0.999, 1e-8, 0.0, self.unc_step, bool(with_uncert_step))]
        if self.peers is not None:
            # reduce-scatter + Adam (all three groups) + all-gather in ONE launch over peer memory (csrc/peer.cu)
            self.peers.adam_step(self.exp_avg, self.exp_avg_sq, groups, self.smooth_total if smooth else None,
                                 keep_grad=[zeroing, zeroing, False]); n += 1
            self._mark('adam')
            return n
        reduce_grads(self.bucket, self.pg)
        p.adam_step_groups(self.theta, self.grad, self.exp_avg, self.exp_avg_sq, groups, zero_grad=True); n += 1
        self._mark('adam')
        return n

    def _mark(self, name):
        if self.stamps is None:
            return
        k = self._stamp_i
        self._stamp_i += 1
        if k >= len(self.stamp_names):
            self.stamp_names.append(name)
        L.check(self.plan.lib.nrt_debug_stamp(self.stamps.data_ptr() + 8 * k, torch.cuda.current_stream().cuda_stream))

    def _smooth_branch(self):
        """Launch the smoothness term on a second stream forked off the current one (inside a capture this becomes a parallel
        branch of the graph): it reads the parameters and rand6 and adds into the table gradient with reductions, so it only
        has to be ordered before the optimiser step."""
        p = self.plan
        if self._branch is None:
            self._branch = torch.cuda.Stream(device=self.dev)
        self._branch.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._branch):
            for _ in range(self.smooth_stagger):
                self.smooth_loss.zero_()
            p.smooth_fwd_bwd(self.P.grid, self.rand6, self.smooth_n, self.smooth_vox, self.smooth_margin, self.smooth_w,
                             self.smooth_loss, self.G.grid, self.ws_smooth, part=self.rank, n_parts=self.world)
        return 2

    def _graph(self, with_uncert_step, smooth=None, host_io=False):
        smooth = self.smooth_on if smooth is None else (bool(smooth) and self.smooth_w > 0)
        key = (bool(with_uncert_step), smooth, bool(host_io))
        if key not in self._graphs:
            # warm-up outside capture (lazy attribute/module init inside the library and torch RNG)
            side = torch.cuda.Stream(device=self.dev)
            saved = [b.clone() for b in self.state.persistent()]
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._body(with_uncert_step, smooth)
            torch.cuda.current_stream().wait_stream(side)
            for b, s in zip(self.state.persistent(), saved):
                b.copy_(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if host_io:          # the iteration's rays from the pinned input buffer ... (memcpy nodes of the same graph)
                    self.inbuf.copy_(self.host_in, non_blocking=True)
                self.launches_per_iter[bool(with_uncert_step)] = self._body(with_uncert_step, smooth)
                if host_io:          # ... and its losses back into pinned host memory
                    self.host_losses.copy_(self.losses, non_blocking=True)
            for b, s in zip(self.state.persistent(), saved):
                b.copy_(s)
            self._graphs[key] = g
        return self._graphs[key]

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

    def step_host(self, packed=None, with_uncert_step=None, smooth=None):
        """One mapping iteration fed from HOST memory and read back to it, as ONE graph launch: H2D of the packed ray batch
        (`self.host_in`, pinned [10*B] = [o | d | rgb | depth]; a `packed` host tensor is first copied into it), the iteration,
        D2H of the losses into `self.host_losses` (pinned [8]; valid once the stream -- or an event recorded after this call --
        has been synchronised).  Against load_packed() + step() + losses.cpu() this saves two API round trips per iteration."""
        if self.host_in is None:
            self.host_in = torch.empty(10 * self.B, dtype=torch.float32, pin_memory=True)
            self.host_losses = torch.zeros(L.N_LOSS, dtype=torch.float32, pin_memory=True)
        if packed is not None and packed.data_ptr() != self.host_in.data_ptr():
            self.host_in.copy_(packed.reshape(-1))
        with_unc = (self.it + 1) % 5 == 0 if with_uncert_step is None else bool(with_uncert_step)
        if self.use_graph:
            self._graph(with_unc, smooth, host_io=True).replay()
        else:
            self.inbuf.copy_(self.host_in, non_blocking=True)
            self.launches_per_iter[with_unc] = self._body(with_unc, smooth)
            self.host_losses.copy_(self.losses, non_blocking=True)
        self.it += 1
        self.state.n_map_steps += 1
        self.state.n_unc_steps += 1 if with_unc else 0
        return self.host_losses

    def step(self, rays_o=None, rays_d=None, target_rgb=None, target_d=None, with_uncert_step=None, smooth=None):
        """One mapping iteration.  Returns the device tensor of the five losses (no host sync).
        with_uncert_step: step + zero the uncertainty-grid Adam group in this iteration (default: every 5th iteration of this
        object, src/slam/coslam/coslam.py:397-399; a caller that follows global_BA's own loop index passes it explicitly);
        smooth: include the smoothness term (default: as constructed; first_frame_mapping runs without it)."""
        if rays_o is not None:
            self.load_rays(rays_o, rays_d, target_rgb, target_d)
        with_unc = (self.it + 1) % 5 == 0 if with_uncert_step is None else bool(with_uncert_step)
        if self.use_graph:
            self._graph(with_unc, smooth).replay()
        else:
            self.launches_per_iter[with_unc] = self._body(with_unc, smooth)
        self.it += 1
        self.state.n_map_steps += 1
        self.state.n_unc_steps += 1 if with_unc else 0
        return self.losses

    def uncert_step(self):
        """A stand-alone uncertainty-grid Adam step + zero_grad (first_frame_mapping steps it once, after all its iterations:
        src/slam/coslam/coslam.py:217-218)."""
        p, o = self.plan, self.n_grid + self.n_dec
        if self.world > 1:            # stand-alone step outside the iteration: plain all-reduce of the accumulated slice
            torch.distributed.all_reduce(self.grad[o:], group=self.pg)
        p.counter_add(self.unc_step, 1)
        p.adam_step(self.theta[o:], self.grad[o:], self.exp_avg[o:], self.exp_avg_sq[o:], 0, 1.0, 0.9, 0.999, 1e-8, 0.0,
                    zero_grad=True, step_dev=self.unc_step)
        self.state.n_unc_steps += 1

    def zero_uncert_grad(self):
        self.grad[self.n_grid + self.n_dec:].zero_()

    def total_loss(self):
        """get_loss_from_ret's scalar (host sync)."""
        l = self.losses[:5].double().cpu()
        w = self.loss_grad.double().cpu()
        tot = float((l * w).sum())
        if self.smooth_on:
            tot += self.smooth_w * float((self.smooth_total if self.peers is not None else self.smooth_loss).item())
        return tot
