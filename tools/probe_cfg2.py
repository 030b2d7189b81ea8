#!/usr/bin/env python
"""BASELINE.json configs[2] shape on one GPU (apartment-sized bound, 2048 rays x 43 samples): the graph-replayed mapping
iteration timed with CUDA events, then each launch of the eager iteration (run under ncu's launch list for the per-kernel split)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from naruto_b200.configs import replica_office0
from naruto_b200.field import FieldPlan, FieldTensors
from naruto_b200.mapper import MappingStep
from naruto_b200.synthetic import SyntheticFrame

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nsd = int(sys.argv[2]) if len(sys.argv) > 2 else 32
apartment = [[-8.0, 8.0], [-6.0, 6.0], [-1.5, 3.5]]
if len(sys.argv) > 3 and sys.argv[3] == 'office0':      # BASELINE.json configs[1]: the headline shape
    from naruto_b200.configs import OFFICE0_BOUND
    apartment = OFFICE0_BOUND
    cfg = replica_office0(n_samples_d=nsd)
else:
    cfg = replica_office0(n_samples_d=nsd, bound=apartment)
plan = FieldPlan(cfg, apartment)
g = torch.Generator().manual_seed(0)
lin = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) / (i ** 0.5)
init = FieldTensors((torch.rand(plan.n_grid_floats, generator=g) * 2 - 1) * 1e-4, lin(32, 80), lin(16, 32), lin(32, 63), lin(3, 32),
                    torch.full(plan.uncert_dims, 3.0))
ms = MappingStep(plan, cfg, B, 'cuda', init=init)
frame = SyntheticFrame(apartment, seed=300)
batches = [frame.sample_packed(B).cuda() for _ in range(4)]
flush = torch.empty(64 * 1024 * 1024, device='cuda')
ts = []
for i in range(30):
    ms.load_packed(batches[i % 4])
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ms.step(); b.record(); b.synchronize()
    ts.append(a.elapsed_time(b))
ts = sorted(ts[6:])
print(f'B={B} S={plan.S} table {plan.n_grid_floats*4/1e6:.1f} MB uncert {tuple(plan.uncert_dims)}: step median {ts[len(ts)//2]*1e3:.1f} us  min {ts[0]*1e3:.1f} us')
# not flushed (what back-to-back iterations see)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(50):
    ms.step()
b.record(); b.synchronize()
print(f'  back to back, no flush: {a.elapsed_time(b)/50*1e3:.1f} us / step')
