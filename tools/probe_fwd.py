#!/usr/bin/env python
"""Experiment driver: CUDA-event time of the training forward (nrt_render_fwd_stats) and of its variants at the bench shape;
NRT_FWD_DEBUG=1 prints the per-phase cycle accounting of sub-CTA 0 of every CTA."""
import ctypes, os, statistics, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_scale_properties import _plan, _rays
from naruto_b200.field import RenderBuffers

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsd = int(sys.argv[2]) if len(sys.argv) > 2 else 117
cfg, plan, P = _plan(nsd, grid_range=1e-4)
o, d, rgb, td = _rays(B, seed=9)
out = RenderBuffers(B, plan.S, 'cuda', per_sample=True, feat=True)
plain = RenderBuffers(B, plan.S, 'cuda', per_sample=False)
stats = plan.new_stats('cuda'); losses = torch.zeros(8, device='cuda')
u = torch.rand(B, plan.S, device='cuda')
step = torch.ones(1, dtype=torch.int32, device='cuda')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
variants = {
    'train+stats, Philox': lambda: plan.render_fwd_stats(P, o, d, rgb, td, out, stats, u=None, seed=5, seed_step=step, losses=losses),
    'train+stats, u given': lambda: plan.render_fwd_stats(P, o, d, rgb, td, out, stats, u=u, losses=losses),
    'train (feat+masks saved), u given, no stats': lambda: plan.render_fwd(P, o, d, td, out, u=u),
    'eval outputs only, Philox': lambda: plan.render_fwd(P, o, d, td, plain, perturb=1, seed=7),
}
for name, fn in variants.items():
    ts = []
    for i in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    print(f'B={B} S={plan.S} {name:48s}: median {statistics.median(ts[2:])*1e3:.1f} us  min {min(ts)*1e3:.1f} us')
    if os.environ.get('NRT_FWD_DEBUG'):
        raw = np.zeros(256 * 8, dtype=np.int64)
        plan.lib.nrt_debug_read(raw.ctypes.data_as(ctypes.c_void_p), -raw.nbytes)
        t = raw.reshape(256, 8)[:148].astype(np.float64)
        for k, nm in enumerate(['prologue', 'ray staging', 'tiles', 'compositing', 'total (sub-CTA 0)', 'total incl. statistics tail']):
            print(f'    {nm:30s} cycles: median {np.median(t[:, k]):9.0f}  max {t[:, k].max():9.0f}')
