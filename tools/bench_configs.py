#!/usr/bin/env python
"""Side measurements for the other BASELINE.json configurations (bench.py itself measures configs[1] and the configs[4] sweep):

  replica43   the reference's shipped shape: office0, 2 148 and 4 096 rays x 43 samples/ray (32 + 11), hash_size 16
  mp3d_large  SURVEY 8(d) config 4: largest shipped MP3D bound, 2^21-entry levels (153.8 MB table: HBM-resident, not
              L2-resident), 192 samples/ray (181 + 11), 8 192 rays

For each: CUDA-event time of one full mapping iteration (CUDA-graph replay, L2 flushed before every step) and of the
forward kernel alone.  Prints one JSON object per configuration.
"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from naruto_b200.configs import MP3D_LARGE_BOUND, OFFICE0_BOUND, mp3d_large, replica_office0   # noqa: E402
from naruto_b200.field import FieldPlan, FieldTensors   # noqa: E402
from naruto_b200.mapper import MappingStep   # noqa: E402
from naruto_b200.synthetic import SyntheticFrame   # noqa: E402


def run(name, cfg, bound, B, steps=20, warmup=5):
    dev = torch.device('cuda:0')
    plan = FieldPlan(cfg, bound)
    g = torch.Generator().manual_seed(0)
    lin = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) / (i ** 0.5)
    init = FieldTensors((torch.rand(plan.n_grid_floats, generator=g) * 2 - 1) * 1e-4, lin(32, 80), lin(16, 32), lin(32, 63),
                        lin(3, 32), torch.full(plan.uncert_dims, 3.0))
    ms = MappingStep(plan, cfg, B, dev, init=init)
    frame = SyntheticFrame(bound, seed=1)
    batches = [frame.sample_packed(B).to(dev) for _ in range(warmup + steps)]
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    ts, tf = [], []
    for i, b in enumerate(batches):
        ms.load_packed(b)
        flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ms.step()
        e1.record()
        e1.synchronize()
        if i >= warmup:
            ts.append(e0.elapsed_time(e1))
    for i in range(6):
        flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.render_fwd(ms.P, ms.rays_o, ms.rays_d, ms.target_d, ms.out, u=ms.u)
        e1.record()
        e1.synchronize()
        tf.append(e0.elapsed_time(e1))
    S = plan.S
    t, f = statistics.median(ts), statistics.median(tf[1:])
    out = {'config': name, 'rays': B, 'samples_per_ray': S, 'table_MB': round(plan.n_grid_floats * 4 / 1e6, 1),
           'step_ms': round(t, 4), 'rays_per_s': round(B / t * 1e3), 'points_per_s': round(B * S / t * 1e3),
           'fwd_ms': round(f, 4), 'fwd_algorithmic_GB_per_s': round(B * (S * 1056 + 60) / f / 1e6, 1),
           'losses_finite': bool(torch.isfinite(ms.losses[:5]).all().item())}
    print(json.dumps(out), flush=True)
    del ms, plan, flush_buf
    torch.cuda.empty_cache()


if __name__ == '__main__':
    run('replica office0 shipped shape (43 samples)', replica_office0(n_samples_d=32), OFFICE0_BOUND, 2148)
    run('replica office0 shipped shape (43 samples)', replica_office0(n_samples_d=32), OFFICE0_BOUND, 4096)
    run('mp3d large, hash_size 21, 192 samples', mp3d_large(), MP3D_LARGE_BOUND, 8192)
