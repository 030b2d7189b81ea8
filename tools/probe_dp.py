#!/usr/bin/env python
"""Where a data-parallel mapping iteration spends its time: globaltimer stamps after every stage of the captured iteration
(NRT_STEP_STAMPS=1, MappingStep._mark) on every rank, printed by rank 0 as the mean over iterations and the max over ranks.  Run under torchrun
(any world size, 1 included):  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/probe_dp.py"""
import os, sys
os.environ['NRT_STEP_STAMPS'] = '1'
os.environ['NRT_PEER_DEBUG'] = '1'
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from naruto_b200.configs import replica_office0, OFFICE0_BOUND
from naruto_b200.field import FieldPlan, FieldTensors
from naruto_b200.mapper import MappingStep
from naruto_b200.synthetic import SyntheticFrame

world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
pg = None
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
    pg = dist.group.WORLD
    dist.all_reduce(torch.zeros(1, device=dev))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = replica_office0(n_samples_d=117)
plan = FieldPlan(cfg, OFFICE0_BOUND)
g = torch.Generator().manual_seed(0)
lin = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) / (i ** 0.5)
init = FieldTensors((torch.rand(plan.n_grid_floats, generator=g) * 2 - 1) * 1e-4, lin(32, 80), lin(16, 32), lin(32, 63), lin(3, 32),
                    torch.full(plan.uncert_dims, 3.0))
ms = MappingStep(plan, cfg, B, dev, init=init, process_group=pg, use_graph=True)
frame = SyntheticFrame(OFFICE0_BOUND, seed=300)          # shards of one global batch: same frame, per-rank pixel draws
frame.gen.manual_seed(3000 + rank)
batches = [frame.sample_packed(B).to(dev) for _ in range(4)]
flush = torch.empty(64 * 1024 * 1024, device=dev)
acc, n = {}, 0
for i in range(40):
    ms.load_packed(batches[i % 4])
    flush.zero_()
    if world > 1:
        dist.barrier()
    ms.step()
    torch.cuda.synchronize()
    if i >= 10 and (i + 1) % 5 != 0:              # (iterations without the uncertainty-grid step: one graph)
        st = ms.stamps.cpu().tolist()
        for k in range(1, len(ms.stamp_names)):
            acc[ms.stamp_names[k]] = acc.get(ms.stamp_names[k], 0.0) + (st[k] - st[k - 1]) * 1e-3
        acc['total'] = acc.get('total', 0.0) + (st[len(ms.stamp_names) - 1] - st[0]) * 1e-3
        if world > 1:                      # phases inside the optimiser launch (csrc/peer.cu, NRT_PEER_DEBUG)
            import ctypes
            tb = (ctypes.c_uint64 * 8)()
            plan.lib.nrt_debug_read(tb, 64 | (1 << 30))
            adam_end = st[len(ms.stamp_names) - 1]
            for nm, a, b in (('  adam: launch->entry', None, 0), ('  adam: grads barrier', 0, 1), ('  adam: slices', 1, 2), ('  adam: fence+count', 2, 3),
                             ('  adam: last CTA wait', 3, 4), ('  adam: done barrier', 4, 5)):
                t0 = st[len(ms.stamp_names) - 2] if a is None else tb[a]
                acc[nm] = acc.get(nm, 0.0) + (tb[b] - t0) * 1e-3
        n += 1
names = list(acc)
t = torch.tensor([acc[k] / n for k in names], dtype=torch.float64, device=dev)
tmax = t.clone()
if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tmean = t.clone(); dist.all_reduce(tmean); tmean /= world
else:
    tmean = t
if rank == 0:
    mc = getattr(ms.state.peers, 'multicast', None) if ms.state.peers is not None else None
    print(f'world {world}, {B} rays/GPU x {plan.S} samples, graph replay, multicast {mc}, us per stage (mean over ranks / max over ranks):')
    for k, a, b in zip(names, tmean.tolist(), tmax.tolist()):
        print(f'  {k:16s} {a:8.1f} {b:8.1f}')
ms.release_graphs()
if world > 1:
    dist.destroy_process_group()
