"""The smoothness term against the reference's OWN CoSLAM.smoothness (third_parties/coslam/coslam.py:245-269), run on the
reference's real Mapper object with its torch.rand draws recorded (oracle/make_golden_smooth.py -> golden/smooth_small.npz):
the oracle restatement on the CPU, the CUDA kernels on the GPU."""
import os

import numpy as np
import pytest
import torch

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'smooth_small.npz'))


def _grid():
    g = torch.Generator().manual_seed(int(GOLD['grid_seed']))
    return (torch.rand(int(GOLD['n_grid']), generator=g) * 2 - 1) * float(GOLD['grid_range'])


def _dense_grad(tag):
    g = torch.zeros(int(GOLD['n_grid']))
    g[torch.from_numpy(GOLD[f'{tag}_grad_idx'].astype(np.int64))] = torch.from_numpy(GOLD[f'{tag}_grad_val'])
    return g


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_oracle_smoothness_matches_the_reference(tag):
    from oracle import naruto_oracle as no
    sp = no.office0_spec()
    sp.smooth_pts = int(GOLD[f'{tag}_pts'])
    P = no.init_params(sp, seed=1)
    P.grid = _grid().requires_grad_(True)
    r6 = torch.from_numpy(GOLD[f'{tag}_rand6'])
    sm = no.smoothness(P, sp, r6[:3], r6[3:].view(1, 1, 1, 3))
    ref = float(GOLD[f'{tag}_loss'])
    assert abs(sm.item() - ref) <= 1e-6 * abs(ref), (sm.item(), ref)
    sm.backward()
    if tag == 'a':
        gref = _dense_grad(tag)
        assert (P.grid.grad - gref).abs().max() <= 1e-6 * gref.abs().max()
    else:
        assert abs(P.grid.grad.double().abs().sum().item() - float(GOLD['b_grad_abs_sum'])) <= 1e-5 * float(GOLD['b_grad_abs_sum'])


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_cuda_smoothness_matches_the_reference(tag):
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.field import FieldPlan
    cfg = replica_office0()
    plan = FieldPlan(cfg, OFFICE0_BOUND)
    n = int(GOLD[f'{tag}_pts'])
    grid = _grid().cuda()
    dgrid = torch.zeros_like(grid)
    loss = torch.zeros(1, device='cuda')
    r6 = torch.from_numpy(GOLD[f'{tag}_rand6']).cuda()
    plan.smooth_fwd_bwd(grid, r6, n, 0.1, 0.05, 1.0, loss, dgrid, plan.smooth_workspace(n, 'cuda'))
    torch.cuda.synchronize()
    ref = float(GOLD[f'{tag}_loss'])
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref), (loss.item(), ref)
    if tag == 'a':
        gref = _dense_grad(tag)
        assert (dgrid.cpu() - gref).abs().max() <= 1e-5 * gref.abs().max()
    else:
        assert abs(dgrid.double().abs().sum().item() - float(GOLD['b_grad_abs_sum'])) <= 1e-4 * float(GOLD['b_grad_abs_sum'])
