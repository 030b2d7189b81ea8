#!/usr/bin/env python
"""Experiment driver: CUDA-event time of nrt_render_bwd (composite_bwd + decode_bwd) at the bench shape, L2 flushed.
NRT_BWD_IMPL / NRT_BWD_DEBUG select variants (read once per process)."""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_scale_properties import _plan, _rays
from naruto_b200.field import FieldTensors, RenderBuffers

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsd = int(sys.argv[2]) if len(sys.argv) > 2 else 117
nogrid = len(sys.argv) > 3 and sys.argv[3] == 'nogrid'
cfg, plan, P = _plan(nsd, grid_range=1e-4)
o, d, rgb, td = _rays(B, seed=9)
out = RenderBuffers(B, plan.S, 'cuda', per_sample=True, feat=True)
stats = plan.new_stats('cuda'); losses = torch.zeros(8, device='cuda')
plan.render_fwd_stats(P, o, d, rgb, td, out, stats, u=None, seed=5, losses=losses)
lg = torch.tensor([5.0, 0.1, 1000.0, 10.0, 0.005], device='cuda')
G = FieldTensors(*[torch.zeros_like(t) for t in P.as_list()])
if nogrid:
    G.grid = None
ws = torch.empty(plan.lib.nrt_render_bwd_workspace(plan.h, B) // 4, device='cuda')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
ts = []
for i in range(12):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.render_bwd(P, o, d, rgb, td, out, stats, lg, G, workspace=ws); b.record(); b.synchronize()
    ts.append(a.elapsed_time(b))
print(f"impl={os.environ.get('NRT_BWD_IMPL','q')} dbg={os.environ.get('NRT_BWD_DEBUG','0')} nogrid={nogrid} B={B} S={plan.S}: "
      f"median {statistics.median(ts[2:])*1e3:.1f} us  min {min(ts)*1e3:.1f} us")
if int(os.environ.get('NRT_BWD_DEBUG', '0')) & 8:
    import ctypes, numpy as np
    raw = np.zeros(256 * 8 + 8 * 32 * 8, dtype=np.int64)
    plan.lib.nrt_debug_read(raw.ctypes.data_as(ctypes.c_void_p), raw.nbytes)
    buf = raw[:2048].reshape(256, 8)
    wb = raw[2048:].reshape(8, 32, 8).astype(np.float64)
    nb = min(148, (B * plan.S + 127) // 128)
    b = buf[:nb].astype(np.float64)
    names = ['prologue', 'mlp_loop(from prologue)', 'scatter_loop(from prologue)', 'to_flush', 'flush']
    d = [b[:, 1] - b[:, 0], b[:, 2] - b[:, 1], b[:, 3] - b[:, 1], b[:, 4] - b[:, 0], b[:, 5] - b[:, 4]]
    for n, v in zip(names, d):
        print(f'  {n:28s} cycles: median {np.median(v):9.0f}  min {v.min():9.0f}  max {v.max():9.0f}')
    print(f'  prologue split: raw loads {np.median(b[:,6]-b[:,0]):.0f}  W23 {np.median(b[:,7]-b[:,6]):.0f}  images {np.median(b[:,1]-b[:,7]):.0f}')
    print(f'  total cycles median {np.median(b[:,5]-b[:,0]):.0f} max {np.max(b[:,5]-b[:,0]):.0f}')
    if os.environ.get('PROBE_WARPS'):
        for cta in (0, 3):
            print(f'  CTA {cta}: warp: loop | wait ready/full | mma | wg | publish | free   (kcycles)')
            for w in range(32):
                r = wb[cta, w] / 1e3
                print(f'    {w:2d} {"scat" if w < 16 else "mlp "} {r[0]:8.1f} | {r[1]:8.1f} | {r[2]:7.1f} | {r[3]:7.1f} | {r[4]:7.1f} | {r[5]:7.1f}')
