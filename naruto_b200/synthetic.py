"""Synthetic Replica-shaped RGB-D frames and ray batches for benchmarks and tests (SURVEY.md 8(d) "synthetic inputs").

A frame is what the reference's simulator hands to CoSLAMNaruto.online_recon_step: colour [H,W,3], depth [H,W] and a
c2w pose; rays come from the pinhole model of get_camera_rays (third_parties/coslam/datasets/utils.py:24-57) with
the floor-divided intrinsics of src/slam/coslam/coslam.py:137-144 (cx=599, cy=339).  Depth is the analytic distance
(in units of the un-normalised ray parameter) to the walls of the scene bound shrunk by 10 %, so every surface lies
inside the bound; 2 % of the pixels get depth 0 (invalid) to exercise the target_d <= 0 branch.
"""
import math

import torch


def camera_rays(H=680, W=1200, fx=600.0, fy=600.0, cx=599.0, cy=339.0):
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing='xy')
    return torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)        # [H,W,3], OpenGL, un-normalised


def random_rotation(gen):
    q = torch.randn(4, generator=gen)
    q = q / q.norm()
    w, x, y, z = q.tolist()
    return torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class SyntheticFrame:
    def __init__(self, bound, seed=0, H=680, W=1200, invalid_frac=0.02):
        g = torch.Generator().manual_seed(seed)
        b = torch.tensor(bound, dtype=torch.float32)
        self.H, self.W = H, W
        self.c2w = torch.eye(4)
        self.c2w[:3, :3] = random_rotation(g)
        self.c2w[:3, 3] = b.mean(1)
        dirs_cam = camera_rays(H, W).reshape(-1, 3)
        # rays_d = sum(d_cam[..., None, :] * R, -1)  (src/slam/coslam/coslam.py:211)
        self.rays_d = torch.sum(dirs_cam[:, None, :] * self.c2w[:3, :3], -1)
        self.rays_o = self.c2w[:3, 3].expand_as(self.rays_d).contiguous()
        lo, hi = b[:, 0] * 0.9 + b.mean(1) * 0.1, b[:, 1] * 0.9 + b.mean(1) * 0.1
        tt = torch.where(self.rays_d > 0, (hi - self.rays_o) / self.rays_d, (lo - self.rays_o) / self.rays_d)
        self.depth = tt.min(dim=1).values.clamp(0.05, 4.9)
        self.depth[torch.rand(H * W, generator=g) < invalid_frac] = 0.0
        self.rgb = torch.rand(H * W, 3, generator=g)
        self.gen = g

    def sample(self, n):
        """n random pixels -> (rays_o[n,3], rays_d[n,3], rgb[n,3], depth[n,1]) on the host."""
        idx = torch.randint(0, self.H * self.W, (n,), generator=self.gen)
        return self.rays_o[idx], self.rays_d[idx], self.rgb[idx], self.depth[idx, None]

    def sample_packed(self, n, pin=False):
        """Same, packed as one [10*n] host buffer [o(3n) | d(3n) | rgb(3n) | depth(n)] for a single H2D copy."""
        o, d, c, z = self.sample(n)
        buf = torch.cat([o.reshape(-1), d.reshape(-1), c.reshape(-1), z.reshape(-1)])
        return buf.pin_memory() if pin else buf
