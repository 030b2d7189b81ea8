"""Writes tests/golden/erp_small.npz from the reference's OWN ERPDepth2Dist (src/layers/erp_conversions.py), imported
unmodified from /root/reference and run on the CPU.  Run in the build container only:  python -m oracle.make_golden_erp"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    sys.path.insert(0, '/root/reference')
    from src.layers.erp_conversions import ERPDepth2Dist
    from src.layers.backprojection import Backprojection
    out = {}
    for tag, s, hw in (('a', 24, (24, 48)), ('b', 32, (40, 64))):
        torch.manual_seed(11 + s)
        m = ERPDepth2Dist(s, hw, torch.device('cpu'))
        H, W = hw
        yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing='ij')
        depth = 1.5 + torch.sin(6 * xx) * torch.cos(4 * yy) + 0.2 * torch.rand(H, W)
        depth[torch.rand(H, W) < 0.02] = 1e8                      # the simulator's "invalid" marker (habitat_simulator.py:142)
        ref = m(depth.reshape(1, 1, H, W))
        rays = torch.matmul(torch.inverse(m.K)[:, :3, :3], Backprojection(s, s).xy)[0]
        out[f'{tag}_s'] = np.int64(s)
        out[f'{tag}_depth'] = depth.numpy()
        out[f'{tag}_dist'] = ref.numpy()
        out[f'{tag}_c2e'] = m.c2e_layer.grid[0, 0].numpy()
        out[f'{tag}_coor'] = torch.cat([l.coor_xy for l in m.e2p_layers]).numpy()
        out[f'{tag}_rays'] = rays.numpy()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'erp_small.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
