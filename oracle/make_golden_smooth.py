"""Writes tests/golden/smooth_small.npz from the reference's OWN `CoSLAM.smoothness` (third_parties/coslam/coslam.py:245-269),
called on the reference's real CoSLAMNaruto object (oracle/ref_harness.build_reference_slam, CPU, tinycudann -> tcnn_shim) with
its two torch.rand draws recorded.  Run in the build container only:  python -m oracle.make_golden_smooth"""
import os

import numpy as np
import torch

from oracle import ref_harness as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRID_SEED, GRID_RANGE = 23, 0.05


def grid_values(n):
    g = torch.Generator().manual_seed(GRID_SEED)
    return (torch.rand(n, generator=g) * 2 - 1) * GRID_RANGE


def main():
    slam, _ = rh.build_reference_slam(None)          # the reference's own model class
    p = slam.model.embed_fn.params
    with torch.no_grad():
        p.copy_(grid_values(p.numel()))
    out = {'grid_seed': np.int64(GRID_SEED), 'grid_range': np.float32(GRID_RANGE), 'n_grid': np.int64(p.numel())}
    real_rand = torch.rand
    for tag, pts in (('a', 12), ('b', 32)):
        draws = []

        def rec(*a, **k):
            v = real_rand(*a, **k)
            draws.append(v.detach().clone().reshape(-1))
            return v

        torch.manual_seed(100 + pts)
        torch.rand = rec
        try:
            loss = slam.smoothness(pts, 0.1, margin=0.05)
        finally:
            torch.rand = real_rand
        assert len(draws) == 2 and draws[0].numel() == 3 and draws[1].numel() == 3
        p.grad = None
        loss.backward()
        out[f'{tag}_pts'] = np.int64(pts)
        out[f'{tag}_rand6'] = torch.cat(draws).numpy()
        out[f'{tag}_loss'] = np.float64(loss.item())
        if pts <= 16:                                  # the gradient itself only for the small lattice (sparse)
            g = p.grad.detach()
            nz = torch.nonzero(g).reshape(-1)
            out[f'{tag}_grad_idx'] = nz.numpy().astype(np.int32)
            out[f'{tag}_grad_val'] = g[nz].numpy()
        else:
            out[f'{tag}_grad_sum'] = np.float64(p.grad.double().sum().item())
            out[f'{tag}_grad_abs_sum'] = np.float64(p.grad.double().abs().sum().item())
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'smooth_small.npz'), **out)
    print({k: (v.shape if hasattr(v, 'shape') and v.shape else v) for k, v in out.items()})


if __name__ == '__main__':
    main()
