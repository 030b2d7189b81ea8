"""ERP depth -> ERP radial distance on the device.

Drop-in for `src.layers.erp_conversions.ERPDepth2Dist` (src/layers/erp_conversions.py:288-354), which the simulator
builds once (`ERPDepth2Dist(512, pano_hw, 'cuda')`, src/simulator/habitat_simulator.py:63) and calls on every step's
panorama depth (`:143`).  Same constructor, same call: `module(erp_depth[1,1,H,W]) -> erp_dist[H,W]`.

The constructor restates the reference's three static look-up grids on the host (they are configuration, computed once):
  * six E2P sampling grids, one per skybox face F R B L U D  (create_erp_coor, src/layers/erp_conversions.py:184-229 over
    src/layers/erp_utils.py:79-108,141-187,225-247,267-286) -- fp32 tensor arithmetic in the reference's op order;
  * the back-projected texel rays K^-1 [u, v, 1]           (Backprojection, src/layers/backprojection.py:31-82);
  * the C2E cube-to-panorama grid                           (C2E.__init__, src/layers/c2e.py:82-130 over
    src/layers/c2e_utils.py:68-93) -- numpy, fp32 angles / fp64 coordinates like the reference.
`forward` is ONE CUDA kernel (csrc/erp.cu) behind `nrt_erp_depth2dist`; there is no torch fallback.
"""
import math

import numpy as np
import torch

from . import _lib as L

_FACE_U_DEG = (0, 90, -180, -90, 0, 0)      # F R B L U D (src/layers/erp_conversions.py:311-312)
_FACE_V_DEG = (0, 0, 0, 0, 90, -90)


def _axis_rotation(angle: torch.Tensor, axis: torch.Tensor) -> torch.Tensor:
    """Rodrigues matrix in fp32, term by term as src/layers/erp_utils.py:79-108 forms it."""
    angle = angle.reshape(1).to(torch.float32)
    axis = axis / torch.sqrt((axis ** 2).sum())
    cos = torch.cos(angle)
    R = torch.diag(cos.repeat(3)) + torch.outer(axis, axis) * (1.0 - cos)
    a = axis * torch.sin(angle)
    a0, a1, a2 = a[0].item(), a[1].item(), a[2].item()
    return R + torch.tensor([[0, -a2, a1], [a2, 0, -a0], [-a1, a0, 0]])


def face_sampling_grid(u_deg: float, v_deg: float, pano_hw, face: int) -> torch.Tensor:
    """[face, face, 2] normalised (x, y) panorama coordinates of one 90-degree skybox face (E2P.coor_xy)."""
    H, W = pano_hw
    fov = torch.tensor(90 * torch.pi / 180)
    yaw, pitch, roll = torch.tensor(-u_deg * torch.pi / 180), torch.tensor(v_deg * torch.pi / 180), torch.tensor(0 * torch.pi / 180)
    half = torch.tan(fov / 2)
    xs = torch.linspace(-half, half, steps=face)
    ys = torch.linspace(-half, half, steps=face)
    pts = torch.ones((face, face, 3))
    pts[:, :, :2] = torch.stack(torch.meshgrid(xs, -ys, indexing='xy'), -1)
    Rx = _axis_rotation(pitch, torch.tensor([1., 0., 0.]))
    Ry = _axis_rotation(yaw, torch.tensor([0., 1., 0.]))
    roll_axis = (torch.tensor([[0., 0., 1.0]]) @ Rx @ Ry)[0]
    Ri = _axis_rotation(roll, roll_axis)
    pts = pts @ Rx @ Ry @ Ri
    x, y, z = pts[:, :, 0], pts[:, :, 1], pts[:, :, 2]
    lon = torch.arctan2(x, z)
    lat = torch.arctan2(y, torch.sqrt(x ** 2 + z ** 2))
    cx = (lon / (2 * torch.pi) + 0.5) * W - 0.5
    cy = (-lat / torch.pi + 0.5) * H - 0.5
    return torch.stack([cx / (W - 1) * 2 - 1, cy / (H - 1) * 2 - 1], dim=-1)


def texel_rays(face: int) -> torch.Tensor:
    """[3, face*face]: K^-1 [u, v, 1] with K[0,0] = K[0,2] = K[1,1] = K[1,2] = face/2 (erp_conversions.py:314-316, 280-283)."""
    K = torch.eye(4)
    K[0, 0] = K[0, 2] = K[1, 1] = K[1, 2] = face / 2
    inv_K = torch.inverse(K.unsqueeze(0))
    grid = np.stack(np.meshgrid(range(face), range(face), indexing='xy'), axis=0).astype(np.float32)
    grid = torch.tensor(grid)
    pix = torch.cat([torch.stack([grid[0].view(-1), grid[1].view(-1)], 0).unsqueeze(0), torch.ones(1, 1, face * face)], 1)
    return torch.matmul(inv_K[:, :3, :3], pix)[0]


def cube_to_pano_grid(face: int, pano_hw) -> torch.Tensor:
    """[H, W, 3] normalised (x, y, face) coordinates into the [6, face, face] cube volume (C2E.grid)."""
    H, W = pano_hw
    lon = np.linspace(-np.pi, np.pi, num=W, dtype=np.float32)
    lat = np.linspace(np.pi, -np.pi, num=H, dtype=np.float32) / 2
    lon, lat = np.meshgrid(lon, lat)
    # face id per pixel (0F 1R 2B 3L 4U 5D): four vertical bands, then the ceiling / floor caps
    tp = np.roll(np.arange(4).repeat(W // 4)[None, :].repeat(H, 0), 3 * W // 8, 1)
    cap = np.zeros((H, W // 4), bool)
    edge = np.linspace(-np.pi, np.pi, W // 4) / 4
    edge = H // 2 - np.round(np.arctan(np.cos(edge)) * H / np.pi).astype(int)
    for col, row in enumerate(edge):
        cap[:row, col] = 1
    cap = np.roll(np.concatenate([cap] * 4, 1), 3 * W // 8, 1)
    tp[cap] = 4
    tp[np.flip(cap, 0)] = 5
    tp = tp.astype(np.int32)
    cx, cy = np.zeros((H, W)), np.zeros((H, W))
    for f in range(4):
        m = tp == f
        cx[m] = 0.5 * np.tan(lon[m] - np.pi * f / 2)
        cy[m] = -0.5 * np.tan(lat[m]) / np.cos(lon[m] - np.pi * f / 2)
    m = tp == 4
    r = 0.5 * np.tan(np.pi / 2 - lat[m])
    cx[m], cy[m] = r * np.sin(lon[m]), r * np.cos(lon[m])
    m = tp == 5
    r = 0.5 * np.tan(np.pi / 2 - np.abs(lat[m]))
    cx[m], cy[m] = r * np.sin(lon[m]), -r * np.cos(lon[m])
    cx = (np.clip(cx, -0.5, 0.5) + 0.5) * (face - 1)
    cy = (np.clip(cy, -0.5, 0.5) + 0.5) * (face - 1)
    g = torch.stack([torch.from_numpy(cx), torch.from_numpy(cy), torch.from_numpy(tp)]).permute(1, 2, 0).contiguous()
    g[..., 2] = g[..., 2] / 5 * 2 - 1
    g[..., 1] = g[..., 1] / (face - 1) * 2 - 1
    g[..., 0] = g[..., 0] / (face - 1) * 2 - 1
    return g.float()


class ERPDepth2Dist(torch.nn.Module):
    """`ERPDepth2Dist(skybox_size, pano_hw, device)`; `forward(erp_depth[1,1,H,W] or [H,W]) -> erp_dist[H,W]`."""

    def __init__(self, skybox_size: int, pano_hw, device, grids=None):
        super().__init__()
        self.skybox_size = int(skybox_size)
        self.pano_hw = (int(pano_hw[0]), int(pano_hw[1]))
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise RuntimeError('naruto_b200.erp.ERPDepth2Dist runs on a CUDA device only (no CPU fallback)')
        self.lib = L.load()
        if grids is None:
            coor = torch.stack([face_sampling_grid(u, v, self.pano_hw, self.skybox_size) for u, v in zip(_FACE_U_DEG, _FACE_V_DEG)])
            grids = (cube_to_pano_grid(self.skybox_size, self.pano_hw), coor, texel_rays(self.skybox_size))
        c2e, coor, rays = grids
        self.register_buffer('c2e_grid', c2e.to(dev, torch.float32).contiguous(), persistent=False)
        self.register_buffer('face_coor', coor.to(dev, torch.float32).contiguous(), persistent=False)
        self.register_buffer('face_rays', rays.to(dev, torch.float32).contiguous(), persistent=False)

    def forward(self, erp_depth: torch.Tensor) -> torch.Tensor:
        H, W = self.pano_hw
        if not erp_depth.is_cuda:
            raise RuntimeError('erp_depth must be a CUDA tensor')
        if erp_depth.numel() != H * W:
            raise ValueError(f'expected one {H}x{W} panorama, got {tuple(erp_depth.shape)}')
        d = erp_depth.reshape(H, W).to(torch.float32).contiguous()
        out = torch.empty(H, W, dtype=torch.float32, device=d.device)
        stream = torch.cuda.current_stream(d.device).cuda_stream
        L.check(self.lib.nrt_erp_depth2dist(L.ptr(d), H, W, L.ptr(self.c2e_grid), L.ptr(self.face_coor), L.ptr(self.face_rays),
                                            self.skybox_size, L.ptr(out), stream))
        return out
