#!/usr/bin/env python
"""BASELINE.json configs[3] shape on one GPU (largest shipped MP3D bound, 2^21-entry levels = 154 MB table that does not fit
L2, 8192 rays x 192 samples): a few training forwards + backwards, for `ncu --set full` captures of the HBM-served gather."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from naruto_b200.configs import MP3D_LARGE_BOUND, mp3d_large
from naruto_b200.field import FieldPlan, FieldTensors, RenderBuffers
from naruto_b200.synthetic import SyntheticFrame

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = mp3d_large()
plan = FieldPlan(cfg, MP3D_LARGE_BOUND)
g = torch.Generator().manual_seed(0)
lin = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / (i ** 0.5)).cuda()
P = FieldTensors(((torch.rand(plan.n_grid_floats, generator=g) * 2 - 1) * 1e-4).cuda(), lin(32, 80), lin(16, 32), lin(32, 63), lin(3, 32),
                 torch.full(plan.uncert_dims, 3.0).cuda())
o, d, rgb, td = [t.cuda() for t in SyntheticFrame(MP3D_LARGE_BOUND, seed=1).sample(B)]
out = RenderBuffers(B, plan.S, 'cuda', per_sample=True, feat=True)
stats, losses = plan.new_stats('cuda'), torch.zeros(8, device='cuda')
lg = torch.tensor([5.0, 0.1, 1000.0, 10.0, 0.005], device='cuda')
G = FieldTensors(*[torch.zeros_like(t) for t in P.as_list()])
ws = torch.empty(plan.lib.nrt_render_bwd_workspace(plan.h, B) // 4, device='cuda')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
for i in range(4):
    flush.zero_()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    plan.render_fwd_stats(P, o, d, rgb, td, out, stats, u=None, seed=5 + i, losses=losses)
    b.record()
    plan.render_bwd(P, o, d, rgb, td, out, stats, lg, G, workspace=ws)
    c.record(); c.synchronize()
    print(f'T=2^21 B={B} S={plan.S}: fwd {a.elapsed_time(b)*1e3:.1f} us  bwd {b.elapsed_time(c)*1e3:.1f} us  table {plan.n_grid_floats*4/1e6:.1f} MB')
