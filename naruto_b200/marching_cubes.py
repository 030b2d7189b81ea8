"""Iso-surface extraction: drop-in for the reference's `marching_cubes` Python module
(third_parties/coslam/external/NumpyMarchingCubes; `import marching_cubes as mcubes` at src/slam/coslam/coslam_utils.py:26,
called at `:145` as `mcubes.marching_cubes(raw.squeeze(), isolevel, truncation=3.0)`).

`marching_cubes(volume, isovalue, truncation)` -> `(vertices float64 [V,3], triangles uint64 [F,3])`, the reference's return
value, equal to it index for index.  `volume` may be a numpy array (any float dtype, like the reference) or a torch tensor that
already lives on the device (the dense sweep of naruto_b200.map_volumes), in which case it never visits the host.  The O(n^3)
part runs in CUDA kernels behind `nrt_mc_extract`; there is no CPU fallback."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def marching_cubes(volume, isovalue, truncation, device='cuda'):
    lib = L.load()
    if isinstance(volume, np.ndarray):
        vol = torch.from_numpy(np.ascontiguousarray(volume, dtype=np.float32))     # the reference reads float(double(value))
    else:
        vol = volume.detach().to(torch.float32)
    if vol.dim() != 3:
        raise RuntimeError('Only three-dimensional arrays are supported.')            # pywrapper.cpp:12
    dev = vol.device if vol.is_cuda else torch.device(device)
    if dev.type != 'cuda':
        raise RuntimeError('naruto_b200.marching_cubes runs on a CUDA device only (no CPU fallback)')
    vol = vol.to(dev).contiguous()
    nx, ny, nz = (int(s) for s in vol.shape)
    ws = torch.empty(max(int(lib.nrt_mc_workspace_bytes(nx, ny, nz)), 16), dtype=torch.uint8, device=dev)
    res = C.c_void_p()
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        L.check(lib.nrt_mc_extract(L.ptr(vol), nx, ny, nz, float(isovalue), float(truncation), L.ptr(ws), stream, C.byref(res)))
    try:
        nv, nf = C.c_int64(), C.c_int64()
        L.check(lib.nrt_mc_result_sizes(res, C.byref(nv), C.byref(nf)))
        verts = np.empty((nv.value, 3), dtype=np.float64)
        faces = np.empty((nf.value, 3), dtype=np.uint64)
        L.check(lib.nrt_mc_result_copy(res, verts.ctypes.data, faces.ctypes.data))
    finally:
        lib.nrt_mc_result_free(res)
    return verts, faces
