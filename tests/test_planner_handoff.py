"""Goal-space uncertainty aggregation: oracle against the reference's own
NarutoPlanner.uncertainty_aggregation_v2 (golden planner_small.npz, CPU); the CUDA kernel against both (GPU)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'planner_small.npz')
CASES = {'a': dict(top_k=400, sub=48), 'b': dict(top_k=4000, sub=300)}


def _case(tag):
    g = np.load(GOLD)
    return {k[2:]: g[k] for k in g.files if k.startswith(tag + '_')}


def _goal_space(dims, zl):
    gx, gy, gz = torch.meshgrid(torch.arange(0, dims[0], 2), torch.arange(0, dims[1], 2), torch.tensor(list(zl)), indexing='ij')
    pts = torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1).float()
    return pts, (gx, gy, gz)


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_oracle_matches_reference_golden(tag):
    from oracle.planner_oracle import goal_aggregate
    c = _case(tag)
    uncert, sdf = torch.from_numpy(c['uncert']), torch.from_numpy(c['sdf'])
    pts, gs = _goal_space(sdf.shape, c['zlevels'])
    coll, aggre = goal_aggregate(uncert, sdf, pts, gs, torch.from_numpy(c['topk']).float(), 0.5 / 0.1, 2 / 0.1, 0.8)
    assert torch.equal(coll, torch.from_numpy(c['coll']))
    assert torch.equal(aggre.reshape(c['aggre'].shape), torch.from_numpy(c['aggre']))


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_kernel_matches_reference_golden(tag):
    from naruto_b200.planner_handoff import GoalSpace
    c = _case(tag)
    gs = GoalSpace(c['sdf'].shape, gs_z_levels=c['zlevels'].tolist(), uncert_top_k=CASES[tag]['top_k'],
                   uncert_top_k_subset=CASES[tag]['sub'])
    ok, out = gs.uncertainty_aggregation_v2([c['uncert'], c['sdf']], topk_vxl=torch.from_numpy(c['topk']).float())
    assert ok
    assert torch.equal(out['gs_uncert_collections'].cpu(), torch.from_numpy(c['coll'])), 'every mask bit-identical'
    assert torch.equal(out['topk_uncert_vxl'].cpu(), torch.from_numpy(c['topk']))
    a, b = out['gs_aggre_uncerts'].cpu(), torch.from_numpy(c['aggre'])
    assert a.shape == b.shape
    assert ((a - b).abs() <= 1e-5 * b.abs() + 1e-7).all()               # summation order of 300 fp32 terms
    assert ((a == 0) == (b == 0)).all()


@pytest.mark.gpu
def test_kernel_matches_oracle_with_own_target_selection():
    """Device volumes in, device selection of the targets, other volume size and z levels; invalid goal space reported."""
    from naruto_b200.planner_handoff import GoalSpace
    from oracle.make_golden_planner import synth_volumes
    from oracle.planner_oracle import goal_aggregate
    dims = (31, 27, 21)
    uncert, sdf = synth_volumes(dims, seed=9)
    gs = GoalSpace(dims, gs_z_levels=None, uncert_top_k_subset=64)
    ud, sd = torch.from_numpy(uncert).cuda(), torch.from_numpy(sdf).cuda()
    ok, out = gs.uncertainty_aggregation_v2([ud, sd])
    assert ok
    topk = out['topk_uncert_vxl'].cpu().float()
    vals = torch.from_numpy(uncert)[topk[:, 0].long(), topk[:, 1].long(), topk[:, 2].long()]
    assert vals.min() >= np.sort(uncert.reshape(-1))[-64], 'targets are the most uncertain voxels'
    gx, gy, gz = torch.meshgrid(gs.gs_x_range, gs.gs_y_range, gs.gs_z_range, indexing='ij')
    coll, aggre = goal_aggregate(torch.from_numpy(uncert), torch.from_numpy(sdf), gs.goal_space_pts.cpu(), (gx, gy, gz), topk,
                                 5.0, 20.0, 0.8)
    assert torch.equal(out['gs_uncert_collections'].cpu(), coll)
    assert torch.allclose(out['gs_aggre_uncerts'].cpu().reshape(-1), aggre, rtol=1e-5, atol=1e-7)
    # solid volume: nothing is safe to go to -> invalid goal space, empty outputs unless forced
    ok2, out2 = gs.uncertainty_aggregation_v2([ud, torch.full_like(sd, -1.0)])
    assert not ok2 and out2 == {}
    ok3, out3 = gs.uncertainty_aggregation_v2([ud, torch.full_like(sd, -1.0)], force_running=True)
    assert ok3 and out3['gs_aggre_uncerts'].abs().max() == 0
