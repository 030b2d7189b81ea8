#!/usr/bin/env python
"""bench.py -- rays/sec of one NARUTO mapping iteration (forward + losses + backward + Adam) at 128 samples/ray.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rays B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one body of the reference's global_BA loop (src/slam/coslam/coslam.py:364-399) on one ray batch of the
BASELINE.json configs[1] workload: Replica office_0 shape, 680x1200 synthetic frame, B rays x 128 samples/ray
(n_samples_d=117 + n_range_d=11), hash_size 16.  Prints ONE JSON line (rank 0).

  value      rays/s with the ray batch already resident in HBM when the timed region starts (CUDA-graph replay)
  e2e        rays/s through the public API with HOST (pinned) ray buffers: H2D of the packed batch and D2H of the
             losses inside the timed region
  roofline   the dominant kernel of the step, timed alone with CUDA events in this process (algorithmic bytes /
             duration vs the measured HBM copy peak)
  cpu_baseline   the oracle (torch restatement of the reference's Python) timed on the host cores, bounded sample

--impl reference times the reference's CPU implementation of the same step (the oracle port; the reference tree and
tinycudann are not on the GPU box) with all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'rays/sec (128 samples/ray) mapping-iter'
UNIT = 'rays/s'
N_SAMPLES_D = 117            # + n_range_d 11 = 128 samples per ray
S = 128
BYTES_PER_RAY_FWD = S * 1056 + 60          # hash gather 1024 B + uncert 32 B per point, + ray I/O
# backward (composite_bwd + fused MLP-backward/scatter): one RMW per gathered table entry and uncertainty corner
# (1024 + 32 B), saved features in (128 B), raw + dL/draw (40 B), z (4 B); feature gradients never leave the SM
BYTES_PER_POINT_BWD = 1024 + 32 + 128 + 40 + 4


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference(n_rays, steps, warmup, seed=0):
    import torch
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the reference arm uses all the host cores it can, whatever launched it
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    from oracle import naruto_oracle as no
    from naruto_b200.synthetic import SyntheticFrame
    spec = no.office0_spec(n_samples_d=N_SAMPLES_D)
    P = no.init_params(spec, seed=seed).clone(requires_grad=True)
    opt = no.MappingOptimisers(P)
    frame = SyntheticFrame(spec.bound, seed=seed)
    times = []
    for it in range(warmup + steps):
        o, d, rgb, td = frame.sample(n_rays)
        t0 = time.perf_counter()
        u = torch.rand(n_rays, spec.n_samples)
        no.mapping_iteration(P, opt, o, d, rgb, td, spec, it, u=u, smooth_draws=(torch.rand(3), torch.rand(1, 1, 1, 3)))
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times), torch.get_num_threads()


def torch_gpu_reference(n_rays, dev, torch, iters=8):
    from oracle import naruto_oracle as no
    from naruto_b200.synthetic import SyntheticFrame
    spec = no.office0_spec(n_samples_d=N_SAMPLES_D)
    P = no.init_params(spec, seed=0).to(dev).clone(requires_grad=True)
    opt = no.MappingOptimisers(P)
    frame = SyntheticFrame(spec.bound, seed=0)
    ts = []
    for it in range(iters):
        o, d, rgb, td = [t.to(dev) for t in frame.sample(n_rays)]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        u = torch.rand(n_rays, spec.n_samples, device=dev)
        no.mapping_iteration(P, opt, o, d, rgb, td, spec, it, u=u,
                             smooth_draws=(torch.rand(3, device=dev), torch.rand(1, 1, 1, 3, device=dev)))
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    t = statistics.median(ts[3:])
    return {'value': n_rays / t * 1e3, 'unit': UNIT, 'ms_per_step': round(t, 3),
            'what': f'oracle/naruto_oracle.py mapping_iteration with torch CUDA ops on the same GPU, {n_rays} rays x {S} samples, fp32'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_rays = args.ref_rays if args.ref_rays > 0 else args.rays
    total, threads = cpu_reference(n_rays, args.steps, args.warmup)
    val = n_rays * args.steps / total
    sample = f'{args.steps} full mapping iterations (fwd+loss+bwd+Adam) of {n_rays} rays x {S} samples on the host'
    line = {
        'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
        'config': {'workload': f'Replica office_0 shape, 680x1200 synthetic frame, mapping iteration, {n_rays} rays x {S} samples/ray '
                               f'(n_samples_d={N_SAMPLES_D}+n_range_d=11), hash_size 16; reference Python path restated in torch '
                               f'(oracle/naruto_oracle.py) on host cores', 'rays_per_step': n_rays},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def time_kernels(plan, ms, torch, flush, reps=10):
    """CUDA-event time of the two big launches of the step, each alone on the current stream, L2 flushed before."""
    B, Sn = ms.B, plan.S
    n_pts = B * Sn
    res = {}

    def timed(fn):
        ts = []
        for _ in range(reps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    # the launch the step itself makes: nrt_render_fwd_stats (features + ReLU masks saved, loss statistics fused, losses finalized)
    res['render_fwd_ws_kernel'] = (timed(lambda: plan.render_fwd_stats(ms.P, ms.rays_o, ms.rays_d, ms.target_rgb, ms.target_d, ms.out,
                                                                       ms.stats, u=None, seed=ms.seed, seed_step=ms.map_step,
                                                                       losses=ms.losses)),
                                   B * BYTES_PER_RAY_FWD)
    gsave = ms.grad.clone()
    t_bwd = timed(lambda: plan.render_bwd(ms.P, ms.rays_o, ms.rays_d, ms.target_rgb, ms.target_d, ms.out, ms.stats, ms.loss_grad,
                                          ms.G, workspace=ms.ws_bwd))
    res['composite_bwd+decode_bwd_q_kernel'] = (t_bwd, n_pts * BYTES_PER_POINT_BWD)
    ms.grad.copy_(gsave)
    return res


def teardown(step_objs):
    """Leave cleanly: CUDA graphs that captured NCCL collectives must be destroyed BEFORE the process group (otherwise
    destroy_process_group() waits forever on the communicator the graphs still reference)."""
    import gc
    import torch
    import torch.distributed as dist
    for m in step_objs:
        m.release_graphs()
    step_objs.clear()
    gc.collect()
    torch.cuda.synchronize()
    if dist.is_available() and dist.is_initialized():
        th = threading.Thread(target=dist.destroy_process_group, daemon=True)
        th.start()
        th.join(20)
        if th.is_alive():                      # never hang the driver: report and leave
            sys.stderr.write('bench.py: destroy_process_group() did not return within 20 s, exiting without it\n')
            sys.stdout.flush()
            os._exit(0)


def side_config(name, cfg, bound, B_local, dev, pg, world, rank, scaling, steps=20, warmup=5):
    """One of the other BASELINE.json configurations: CUDA-event time of the full mapping iteration (graph replay, L2 flushed
    before every step, max over ranks) and of the forward launch alone on rank 0."""
    import torch
    import torch.distributed as dist
    from naruto_b200.field import FieldPlan, FieldTensors
    from naruto_b200.mapper import MappingStep
    from naruto_b200.synthetic import SyntheticFrame
    plan = FieldPlan(cfg, bound)
    g = torch.Generator().manual_seed(0)
    lin = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) / (i ** 0.5)
    init = FieldTensors((torch.rand(plan.n_grid_floats, generator=g) * 2 - 1) * 1e-4, lin(32, 80), lin(16, 32), lin(32, 63),
                        lin(3, 32), torch.full(plan.uncert_dims, 3.0))
    ms = MappingStep(plan, cfg, B_local, dev, init=init, process_group=pg)
    del init
    frame = SyntheticFrame(bound, seed=300)            # one frame, every rank its own pixel draws (shards of one global batch)
    frame.gen.manual_seed(3000 + rank)
    batches = [frame.sample_packed(B_local).to(dev) for _ in range(4)]
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    evs = []
    for i in range(warmup + steps):
        ms.load_packed(batches[i % 4])
        flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ms.step()
        b.record()
        if i == warmup - 1:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
        if i >= warmup:
            evs.append((a, b))
    torch.cuda.synchronize()
    t = sum(a.elapsed_time(b) for a, b in evs) / 1e3
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = tt.item()
    Sn = plan.S
    out = {'workload': name, 'rays_per_step_per_gpu': B_local, 'rays_per_step': B_local * world, 'samples_per_ray': Sn,
           'table_MB': round(plan.n_grid_floats * 4 / 1e6, 1), 'scaling': scaling, 'steps': steps,
           'ms_per_step': round(1e3 * t / steps, 4), 'rays_per_s': B_local * world * steps / t,
           'losses_finite': bool(torch.isfinite(ms.losses[:5]).all().item())}
    if rank == 0:
        tf = []
        for i in range(6):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            plan.render_fwd_stats(ms.P, ms.rays_o, ms.rays_d, ms.target_rgb, ms.target_d, ms.out, ms.stats, u=None, seed=ms.seed,
                                  seed_step=ms.map_step, losses=ms.losses)
            b.record()
            b.synchronize()
            tf.append(a.elapsed_time(b))
        f = statistics.median(tf[1:])
        hbm, _ = peaks()
        gbs = B_local * (Sn * 1056 + 60) / f / 1e6
        out.update({'fwd_ms': round(f, 4), 'fwd_algorithmic_GB_per_s': round(gbs, 1), 'fwd_frac_of_hbm_peak': round(gbs / hbm, 4)})
    ms.release_graphs()
    del ms, plan, flush_buf, batches
    torch.cuda.empty_cache()
    return out


def dropin_path(cfg, bound, init, host_batches, B, dev, torch, flush, steps, warmup):
    """The import-swap path of INTEGRATION.md section 1, timed: what the reference's global_BA loop body
    (src/slam/coslam/coslam.py:364-399) does per iteration once `JointEncoding` is naruto_b200's class -- model.forward ->
    get_loss_from_ret(smooth=True) (its smoothness() goes through model.query_sdf(embed=True), tp/coslam.py:245-269) ->
    loss.backward() -> torch.optim.Adam.step() / zero_grad() (+ the every-5th uncert_optim.step()), with HOST ray buffers and
    a D2H read of the loss inside the timed region.  This harness stands in for the reference's caller (the reference tree
    is not on the GPU box); every tensor op of the path itself runs in libnaruto_b200.so behind the drop-in class."""
    from naruto_b200.scene_rep import JointEncodingNaruto
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):       # the class prints 'SDF resolution' like the reference's; stdout carries ONE JSON line
        m = JointEncodingNaruto(cfg, torch.tensor(bound)).to(dev)
    unc_opt = torch.optim.Adam(params=[m.get_uncert_grid(0.1)], lr=1)
    with torch.no_grad():
        for dst, src in zip(m._tensors().as_list(), init.as_list()):
            dst.copy_(src.to(dev))
    map_opt = torch.optim.Adam([{'params': m.decoder.parameters(), 'weight_decay': 1e-6, 'lr': cfg['mapping']['lr_decoder']},
                                {'params': m.embed_fn.parameters(), 'eps': 1e-15, 'lr': cfg['mapping']['lr_embed']}], betas=(0.9, 0.99))
    t = cfg['training']
    bb = torch.tensor(bound, dtype=torch.float32, device=dev)
    n = t['smooth_pts']
    r = torch.arange(0, n - 1, device=dev)
    lattice = torch.stack(torch.meshgrid(r, r, r, indexing='ij'), dim=-1).float()

    def smoothness():
        off_max = bb[:, 1] - bb[:, 0] - (n - 1) * t['smooth_vox'] - 2 * t['smooth_margin']
        off = torch.rand(3, device=dev) * off_max + t['smooth_margin']
        pts = (lattice + torch.rand((1, 1, 1, 3), device=dev)) * t['smooth_vox'] + bb[:, 0] + off
        f = m.query_sdf((pts - bb[:, 0]) / (bb[:, 1] - bb[:, 0]), embed=True)
        tv = ((f[1:] - f[:-1]) ** 2).sum() + ((f[:, 1:] - f[:, :-1]) ** 2).sum() + ((f[:, :, 1:] - f[:, :, :-1]) ** 2).sum()
        return tv / n ** 3

    m.train()
    map_opt.zero_grad()
    unc_opt.zero_grad()
    evs = []
    for i in range(warmup + steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        hb = host_batches[i % len(host_batches)].to(dev, non_blocking=True)
        o, d, rgb, td = hb[0:3 * B].view(B, 3), hb[3 * B:6 * B].view(B, 3), hb[6 * B:9 * B].view(B, 3), hb[9 * B:10 * B].view(B, 1)
        ret = m.forward(o, d, rgb, td)
        loss = (t['rgb_weight'] * ret['rgb_loss'] + t['depth_weight'] * ret['depth_loss'] + t['sdf_weight'] * ret['sdf_loss']
                + t['fs_weight'] * ret['fs_loss'] + t['smooth_weight'] * smoothness() + t['uncert_weight'] * ret['uncert_loss'])
        loss.backward(retain_graph=True)
        map_opt.step()
        map_opt.zero_grad()
        if (i + 1) % 5 == 0:
            unc_opt.step()
            unc_opt.zero_grad()
        host_loss = loss.item()                      # D2H of the step's result
        b.record()
        if i >= warmup:
            evs.append((a, b))
    torch.cuda.synchronize()
    tsec = sum(a.elapsed_time(b) for a, b in evs) / 1e3
    return {'value': B * steps / tsec, 'unit': UNIT, 'ms_per_step': round(1e3 * tsec / steps, 4), 'steps': steps,
            'h2d_bytes_per_step': 10 * B * 4, 'd2h_bytes_per_step': 4, 'finite': bool(host_loss == host_loss),
            'what': 'JointEncodingNaruto.forward -> get_loss_from_ret(smooth=True) -> loss.backward() -> torch.optim.Adam, eager '
                    '(the one-import swap; no CUDA graph, Python optimiser)'}


def mapper_path(dev, torch, n_calls=6, n_keyframes=24):
    """The Mapper-level drop-in, timed end to end: `global_BA(slam, batch, cur_frame_id)` of
    naruto_b200/coslam_mapper.py = src/slam/coslam/coslam.py:246-407 as shipped for Replica office_0 (mapping.sample 2048,
    iters 10, 32+11 samples/ray, active ray sampling, filter_depth) on a duck-typed CoSLAMNaruto whose key-frame database
    holds `n_keyframes` 680x1200 frames (past 20 key frames the batch size no longer changes from call to call: the steady
    state of a run; before that every new size runs un-captured, see FusedMapper.step_for).  `batch` is what the reference's data loader hands over: HOST tensors; the timed
    region of a call holds the H2D copy of the frame, the per-iteration device-side sampling (key-frame + current-frame draws,
    pose application, uncertainty-guided selection), 10 fused mapping iterations and the D2H read of the loss."""
    import contextlib
    import io
    import types
    from naruto_b200 import coslam_mapper as cm
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.ray_sampler import DeviceActiveRaySampler, DeviceKeyFrameDatabase
    from naruto_b200.scene_rep import JointEncodingNaruto
    from naruto_b200.synthetic import SyntheticFrame, camera_rays
    cfg = replica_office0()
    cfg['mapping']['active_ray'] = True
    H, W = 680, 1200
    with contextlib.redirect_stdout(io.StringIO()):
        m = JointEncodingNaruto(cfg, torch.tensor(OFFICE0_BOUND)).to(dev)
    map_opt = torch.optim.Adam([{'params': m.decoder.parameters(), 'weight_decay': 1e-6, 'lr': cfg['mapping']['lr_decoder']},
                                {'params': m.embed_fn.parameters(), 'eps': 1e-15, 'lr': cfg['mapping']['lr_embed']}], betas=(0.9, 0.99))
    unc_opt = torch.optim.Adam(params=[m.get_uncert_grid(0.1)], lr=1)
    s = types.SimpleNamespace(config=cfg, model=m, map_optimizer=map_opt, uncert_optim=unc_opt, device=dev,
                              dataset=types.SimpleNamespace(H=H, W=W), est_c2w_data={}, est_c2w_data_rel={}, step=0,
                              info_printer=lambda *a, **k: None)
    n_save = int(H * W * cfg['mapping']['n_pixels'])
    every = cfg['mapping']['keyframe_every']
    s.keyframeDatabase = DeviceKeyFrameDatabase(cfg, H, W, n_keyframes + n_calls + 4, n_save, dev)
    s.active_ray_sampler = DeviceActiveRaySampler(config=cfg, num_uncert_sample=500, oversample_mul=4)
    s.cached_uncert = torch.rand(*m.plan.uncert_dims, device=dev)
    dirs = camera_rays(H, W)

    def frame(fid):
        f = SyntheticFrame(OFFICE0_BOUND, seed=900 + fid, H=H, W=W)
        s.est_c2w_data[fid] = f.c2w.to(dev)
        return {'frame_id': torch.tensor([fid]), 'c2w': f.c2w.unsqueeze(0), 'rgb': f.rgb.reshape(1, H, W, 3),
                'depth': f.depth.reshape(1, H, W), 'direction': dirs.unsqueeze(0)}

    for k in range(n_keyframes):
        s.keyframeDatabase.add_keyframe(frame(k * every), filter_depth=cfg['mapping']['filter_depth'])
    ts, loss = [], None
    iters = cfg['mapping']['iters']
    for c in range(n_calls + 3):
        fid = (n_keyframes + c) * every
        batch = frame(fid)
        torch.cuda.synchronize()
        prof = None
        if os.environ.get('NRT_PROFILE_MAPPER') and c == n_calls + 2:      # host-side profile of the last call -> stderr
            import cProfile
            prof = cProfile.Profile()
            prof.enable()
        t0 = time.perf_counter()
        out = cm.global_BA(s, batch, fid)
        loss = out[1].item()                        # D2H of the call's result
        ts.append(time.perf_counter() - t0)
        if prof is not None:
            import pstats
            prof.disable()
            pstats.Stats(prof, stream=sys.stderr).sort_stats('cumulative').print_stats(45)
        s.keyframeDatabase.add_keyframe(batch, filter_depth=cfg['mapping']['filter_depth'])
    ts = ts[3:]                                     # the first call of a batch size runs eagerly, the second builds the graphs
    tsec = sum(ts) / len(ts)
    fm = cm._mapper(s)
    rays = max(fm.steps) if fm.steps else cfg['mapping']['sample']       # mapping.sample + the current-frame tail
    fm.release()
    return {'what': 'coslam_mapper.global_BA on HOST frames: H2D of the 680x1200 colour + depth images + device-side sampling + '
                    f'{iters} fused mapping iterations + D2H of the loss, per call (wall clock, synchronised)',
            'ms_per_call': round(1e3 * tsec, 3), 'iterations_per_call': iters, 'rays_per_iteration': rays,
            'key_frames': n_keyframes, 'value': rays * iters / tsec, 'unit': UNIT, 'finite': bool(loss == loss),
            'h2d_bytes_per_call': H * W * 4 * 4, 'd2h_bytes_per_call': 4,
            'note': 'rgb + depth go up with every call; the camera-ray image (the same host tensor for every frame of a run) is '
                    'uploaded once (ray_sampler.device_directions)'}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.field import FieldPlan, FieldTensors
    from naruto_b200.mapper import MappingStep
    from naruto_b200.synthetic import SyntheticFrame

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    pg = None
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            pg = dist.group.WORLD
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B, K, W = args.rays, args.steps, args.warmup

    cfg = replica_office0(n_samples_d=N_SAMPLES_D)
    plan = FieldPlan(cfg, OFFICE0_BOUND)
    assert plan.S == S
    # random-init weights of the reference architecture: tcnn grid init U(-1e-4,1e-4), torch Linear default init
    g = torch.Generator().manual_seed(0)
    lin = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) / (i ** 0.5)
    init = FieldTensors((torch.rand(plan.n_grid_floats, generator=g) * 2 - 1) * 1e-4, lin(32, 80), lin(16, 32), lin(32, 63),
                        lin(3, 32), torch.full(plan.uncert_dims, 3.0))
    ms = MappingStep(plan, cfg, B, dev, init=init, process_group=pg, use_graph=not args.no_graph)
    # data parallel = one global batch of world x B rays split evenly: every rank draws its own pixels of the SAME synthetic frame
    # (statistically identical shards, like shards of a batch sampled from one key-frame database; a different camera pose per
    # rank would add a 5-8 % spread of the forward time between ranks that a sharded batch does not have)
    frame = SyntheticFrame(OFFICE0_BOUND, seed=100)
    frame.gen.manual_seed(1000 + rank)
    host_batches = [frame.sample_packed(B, pin=True) for _ in range(W + K)]
    dev_batches = [hb.to(dev) for hb in host_batches]
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # 256 MB > 126 MB L2

    def flush():
        flush_buf.zero_()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run_phase(host_inputs):
        """W warm-up + K timed steps.  Each timed step: L2 flush (untimed), then [start] inputs -> step -> result [end]."""
        evs = []
        for i in range(W + K):
            if host_inputs:
                # the step's rays are in the iteration's pinned HOST input buffer when the timed region starts; the region holds
                # the H2D copy of them, the iteration and the D2H copy of its losses into pinned host memory (one graph launch:
                # MappingStep.step_host), and the host waits for the result before it prepares the next step
                ms.host_in.copy_(host_batches[i])
                flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                host_losses = ms.step_host()
                b.record()
                b.synchronize()
                e2e_check.append(float(host_losses[0]))
            else:
                ms.load_packed(dev_batches[i])                  # already resident in HBM: staged before the timed region
                flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ms.step()
                b.record()
            if i == W - 1:
                barrier()
            if i >= W:
                evs.append((a, b))
        barrier()
        t = sum(a.elapsed_time(b) for a, b in evs) / 1e3
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = tt.item()
        return t

    # graph capture + lazy init outside any timing
    ms.load_packed(dev_batches[0])
    for _ in range(5):
        ms.step()
    e2e_check = []
    for _ in range(5):
        ms.step_host(host_batches[0])
    barrier()
    it0 = ms.it

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_dev = run_phase(host_inputs=False)
    t_e2e = run_phase(host_inputs=True)
    clocks = sampler.stop() if rank == 0 else None

    launches = sum(ms.launches_per_iter[(it0 + W + i + 1) % 5 == 0] for i in range(K))
    finite = bool(torch.isfinite(ms.losses[:5]).all().item()) and all(v == v for v in e2e_check)
    value = world * B * K / t_dev
    e2e = world * B * K / t_e2e

    roof, kern, cpu = None, None, None
    if rank == 0:
        hbm, how = peaks()
        kt = time_kernels(plan, ms, torch, flush)
        kern = {k: {'ms': round(v[0], 4), 'algorithmic_GB_per_s': round(v[1] / v[0] / 1e6, 1)} for k, v in kt.items()}
        top = max(kt, key=lambda k: kt[k][0])
        ach = kt[top][1] / kt[top][0] / 1e6
        traffic, lsu, tjd = None, None, {}
        tj = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tj):
            # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (same workload)
            tjd = json.load(open(tj))
            tr = tjd['dram_bytes_per_launch']
            traffic = tr.get('decode_bwd_q_kernel' if 'bwd' in top else 'render_fwd_ws_kernel')
            lsu = tjd.get('decode_bwd_q_kernel_lsu') if 'bwd' in top else None
        roof = {'bound': 'hbm', 'kernel': top, 'achieved': round(ach, 1), 'peak': hbm, 'unit': 'GB/s', 'frac': round(ach / hbm, 4),
                'traffic': traffic, 'peak_source': how,
                'note': 'algorithmic bytes (SURVEY 8d) / CUDA-event time of the kernel launched alone, L2 flushed; at hash_size 16 '
                        'the 6.5 MB table is L2-resident so the gather fraction is an accounting convention; the backward entry '
                        'is the launches of nrt_render_bwd: composite_bwd_kernel (7% of it) + decode_bwd_q_kernel + wgrad_reduce_kernel; traffic is '
                        'decode_bwd_q_kernel\'s',
                'lsu_floor': None if not lsu else {
                    'what': 'what actually bounds decode_bwd_q: the SM load/store pipe.  A no-return reduction costs 0.79 ns per lane per SM '
                            'whatever its width (tools/micro/red_throughput.cu); lanes / shared wavefronts / shuffles from the committed ncu capture',
                    'red_lane_floor_ms': round(lsu['red_lanes_per_launch'] * lsu['red_ns_per_lane_per_sm'] * 1e-6 / 148, 4),
                    'red_lanes_per_launch': lsu['red_lanes_per_launch'], 'shared_wavefronts_per_launch': lsu['shared_wavefronts_per_launch']},
                'hash_gather': {'kernel': 'render_fwd_ws_kernel', 'achieved': kern['render_fwd_ws_kernel']['algorithmic_GB_per_s'],
                                'frac': round(kern['render_fwd_ws_kernel']['algorithmic_GB_per_s'] / hbm, 4),
                                # what bounds the gather is the rate at which an SM's load pipe serves divergent 8-byte reads:
                                # corner reads per second against the measured rate of independent random reads from an
                                # L2-resident table (tools/micro/random_gather.cu; x-neighbour pairs share a wavefront, so > 1 is possible)
                                'corner_reads_G_per_s': round(B * S * 128 / kt['render_fwd_ws_kernel'][0] / 1e6, 1),
                                'random_read_ceiling_G_per_s': tjd.get('random_gather_microbench', {}).get('l2_resident_0p5MB_G_reads_per_s')}}
    sweep = None
    if rank == 0 and world == 1 and args.sweep_rays > 0:
        # BASELINE.json configs[4]: uncertainty-only forward sweep, 1M rays x 128 samples, no backward, only the per-ray
        # depth / colour / uncertainty materialised (no raw / z_vals / saved features)
        from naruto_b200.field import RenderBuffers
        n_sw = args.sweep_rays
        o, d, _, td = SyntheticFrame(OFFICE0_BOUND, seed=7).sample(n_sw)
        o, d, td = o.to(dev), d.to(dev), td.to(dev)
        sw_out = RenderBuffers(n_sw, S, dev, per_sample=False)
        ts = []
        for i in range(4):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            plan.render_fwd(ms.P, o, d, td, sw_out, perturb=1, seed=1234 + i)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        t_sw = statistics.median(ts[1:])
        hbm_sw, _ = peaks()
        gbs = n_sw * BYTES_PER_RAY_FWD / t_sw / 1e6
        sweep = {'workload': f'forward-only sweep, {n_sw} rays x {S} samples, outputs rgb/depth/uncert only, in-kernel Philox jitter',
                 'ms': round(t_sw, 3), 'rays_per_s': n_sw / t_sw * 1e3, 'algorithmic_GB_per_s': round(gbs, 1),
                 'frac_of_hbm_peak': round(gbs / hbm_sw, 4), 'finite': bool(torch.isfinite(sw_out.depth).all().item())}
        del sw_out, o, d, td
    gpu_torch = None
    if rank == 0 and world == 1 and args.torch_gpu_baseline:
        # the reference's Python path restated in torch ops (the oracle), run on THIS GPU -- the
        # "reference PyTorch-GPU path" denominator of the north_star's >= 10x target.  ~150 ATen launches per iteration.
        try:
            gpu_torch = torch_gpu_reference(B, dev, torch)
        except Exception as e:          # e.g. out of memory: report, never fail the bench
            gpu_torch = {'error': repr(e)[:200]}
    dropin = None
    if rank == 0 and world == 1 and args.dropin:
        try:
            dropin = dropin_path(cfg, OFFICE0_BOUND, init, host_batches, B, dev, torch, flush, min(K, 20), 5)
        except Exception as e:
            dropin = {'error': repr(e)[:300]}
    mapper = None
    if rank == 0 and world == 1 and args.dropin:
        try:
            mapper = mapper_path(dev, torch)
        except Exception as e:
            mapper = {'error': repr(e)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu, k_cpu = 1024, 5
        tot, threads = cpu_reference(n_cpu, k_cpu, 1)
        cpu = {'value': n_cpu * k_cpu / tot, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': f'{k_cpu} mapping iterations of {n_cpu} rays x {S} samples (oracle/naruto_oracle.py, torch CPU, '
                         f'{os.cpu_count()} host cpus)'}
    side = None
    if args.side_configs:
        # BASELINE.json configs[2] (2 048-ray global batches split over the ranks: STRONG scaling; synthetic apartment-sized
        # bound) and configs[3] (largest shipped MP3D bound, 2^21-entry levels = 154 MB table that does not
        # fit L2, 192 samples/ray, 8 192 rays per GPU).  Every rank takes part (the iteration holds collectives).
        from naruto_b200.configs import MP3D_LARGE_BOUND, mp3d_large
        apartment = [[-8.0, 8.0], [-6.0, 6.0], [-1.5, 3.5]]
        side = {}
        if 2048 % world == 0:
            side['apartment0_2048ray_batches'] = side_config(
                'synthetic apartment_0-sized bound, 2048-ray GLOBAL batch split over the ranks, 43 samples/ray (32+11), hash_size 16',
                replica_office0(n_samples_d=32, bound=apartment), apartment, 2048 // world, dev, pg, world, rank, 'strong')
        side['mp3d_large_hash21_192samples'] = side_config(
            'MP3D YmJkqBEsHnH bound, 2^21-entry levels (HBM-resident table), 8192 rays/GPU x 192 samples/ray (181+11)',
            mp3d_large(), MP3D_LARGE_BOUND, 8192, dev, pg, world, rank, 'weak')
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    if rank != 0:
        teardown([ms])
        return
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': 1e3 * t_dev / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': f'Replica office_0 shape, 680x1200 synthetic frame, mapping iteration (render fwd + losses + bwd + '
                               f'smoothness + Adam), {B} rays/GPU x {S} samples/ray (n_samples_d={N_SAMPLES_D}+n_range_d=11), '
                               f'hash_size 16, random-init weights',
                   'rays_per_step_per_gpu': B, 'samples_per_ray': S, 'parallelism': f'ray-sharded dp{world} (shards = per-rank pixel draws of one frame), '
                                  + ('exchanges over NVLink peer memory inside our kernels' if ms.peers is not None else
                                     'NCCL all-reduces' if world > 1 else 'single shard'),
                   'l2': 'explicit 256 MB L2 flush before every timed step', 'cuda_graph': not args.no_graph},
        'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': 10 * B * 4, 'd2h_bytes_per_step': 8 * 4,
                'ms_per_step': 1e3 * t_e2e / K},
        'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'kernels': kern, 'losses_finite': finite,
    }
    if sweep is not None:
        line['sweep'] = sweep
    if gpu_torch is not None:
        line['torch_gpu_baseline'] = gpu_torch
    if cpu is not None:
        line['cpu_baseline'] = cpu
    if dropin is not None:
        line['e2e_dropin'] = dropin
    if mapper is not None:
        line['e2e_mapper'] = mapper
    if side is not None:
        line['configs'] = side
    print(json.dumps(line))
    sys.stdout.flush()
    teardown([ms])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--rays', type=int, default=4096, help='rays per GPU per mapping iteration')
    ap.add_argument('--ref-rays', type=int, default=0, help='rays per step for --impl reference (0 = the same as --rays)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--sweep-rays', type=int, default=1 << 20, help='rays of the forward-only sweep (0 = skip)')
    ap.add_argument('--no-torch-gpu-baseline', dest='torch_gpu_baseline', action='store_false',
                    help='skip timing the oracle (reference Python path restated in torch ops) on the GPU')
    ap.add_argument('--no-side-configs', dest='side_configs', action='store_false',
                    help='skip the BASELINE.json configs[2] / configs[3] side measurements')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-dropin', dest='dropin', action='store_false', help='skip timing the import-swap (autograd) path')
    ap.set_defaults(torch_gpu_baseline=True, side_configs=True, dropin=True)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        if args.steps > 10:
            args.steps = 10          # bounded sample: ~1.5 s of host work per 4096-ray iteration
        if args.warmup > 3:
            args.warmup = 3
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
