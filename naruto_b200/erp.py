"""ERP depth -> ERP radial distance on the device.

Drop-in for `src.layers.erp_conversions.ERPDepth2Dist` (src/layers/erp_conversions.py:288-354), which the simulator
builds once (`ERPDepth2Dist(512, pano_hw, 'cuda')`, src/simulator/habitat_simulator.py:63) and calls on every step's
panorama depth (`:143`).  Same constructor, same call: `module(erp_depth[1,1,H,W]) -> erp_dist[H,W]`.

Nothing is tabulated: the reference's constructor builds three static look-up grids (six E2P sampling grids, the
back-projected texel rays, the C2E cube-to-panorama grid: 46 MB at the simulator's size); here every entry is a closed-form
function of the output pixel evaluated inside the kernel (csrc/erp.cu: erp_depth2dist_analytic_kernel), and the host passes
only the six 3x3 face frames.  `grids=(c2e, coor, rays)` -- e.g. the tensors of the reference's own constructor -- selects
the grid-fed kernel instead (bit-identical texel choices by construction; used by the parity tests).
`forward` is ONE CUDA kernel behind `nrt_erp_depth2dist_analytic` / `nrt_erp_depth2dist`; there is no torch fallback.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L

_FACE_U_DEG = (0, 90, -180, -90, 0, 0)      # F R B L U D: the six 90-degree views of the skybox
_FACE_V_DEG = (0, 0, 0, 0, 90, -90)


def face_frames(u_deg=_FACE_U_DEG, v_deg=_FACE_V_DEG) -> np.ndarray:
    """[6, 3, 3] float32 frames M_f with  p_panorama = p_face @ M_f  for a tangent-plane point p_face = [x, -y, 1].

    A view looking `u` degrees to the right and `v` degrees up is the pitch about x followed by the yaw about y (angle -u),
    both taken as ordinary right-handed rotation matrices and applied to ROW vectors.  Evaluated in float64; entries of a
    right-angle view are snapped to the exact -1 / 0 / 1 they stand for (the reference forms the same frames in fp32
    through Rodrigues' formula and carries cos(pi/2) ~ 4e-8 residues, a 1e-7-pixel effect)."""
    out = np.zeros((len(u_deg), 3, 3))
    for f, (u, v) in enumerate(zip(u_deg, v_deg)):
        yaw, pitch = math.radians(-u), math.radians(v)
        cp, sp, cy, sy = math.cos(pitch), math.sin(pitch), math.cos(yaw), math.sin(yaw)
        about_x = np.array([[1.0, 0.0, 0.0], [0.0, cp, -sp], [0.0, sp, cp]])
        about_y = np.array([[cy, 0.0, sy], [0.0, 1.0, 0.0], [-sy, 0.0, cy]])
        m = about_x @ about_y
        snapped = np.round(m)
        out[f] = np.where(np.abs(m - snapped) < 1e-12, snapped, m)
    return out.astype(np.float32)


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
        H, W = self.pano_hw
        if W % 4 != 0 or W < 8 or H < 2 or self.skybox_size < 2:
            raise ValueError('ERPDepth2Dist: panorama width must be a multiple of 4 (four face bands of W/4 columns), height >= 2, skybox >= 2')
        self.grids = None
        if grids is not None:
            c2e, coor, rays = grids
            self.grids = tuple(t.to(dev, torch.float32).contiguous() for t in (c2e, coor, rays))
        self.frames = np.ascontiguousarray(face_frames().reshape(-1))
        # tan(fov / 2) of a 90-degree face, in fp32 like the reference's tensor ops
        self.x_max = float(np.tan(np.float32(90 * np.pi / 180) / np.float32(2), dtype=np.float32))

    def forward(self, erp_depth: torch.Tensor) -> torch.Tensor:
        H, W = self.pano_hw
        if not erp_depth.is_cuda:
            raise RuntimeError('erp_depth must be a CUDA tensor')
        if erp_depth.numel() != H * W:
            raise ValueError(f'expected one {H}x{W} panorama, got {tuple(erp_depth.shape)}')
        d = erp_depth.reshape(H, W).to(torch.float32).contiguous()
        out = torch.empty(H, W, dtype=torch.float32, device=d.device)
        stream = torch.cuda.current_stream(d.device).cuda_stream
        if self.grids is not None:
            c2e, coor, rays = self.grids
            L.check(self.lib.nrt_erp_depth2dist(L.ptr(d), H, W, L.ptr(c2e), L.ptr(coor), L.ptr(rays), self.skybox_size, L.ptr(out),
                                                stream))
        else:
            L.check(self.lib.nrt_erp_depth2dist_analytic(L.ptr(d), H, W, self.skybox_size,
                                                         self.frames.ctypes.data_as(C.POINTER(C.c_float)), self.x_max, L.ptr(out),
                                                         stream))
        return out
