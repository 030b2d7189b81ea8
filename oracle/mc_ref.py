"""TEST INFRASTRUCTURE -- ctypes access to oracle/_ref/libmc_ref.so: the reference's own marching_cubes.cpp compiled where it
lies (oracle/build_ref.py).  `marching_cubes(volume, isovalue, truncation)` has the signature and return value of the
reference's `mcubes.marching_cubes` (marching_cubes/src/_mcubes.pyx:20-25): vertices float64 [V,3], faces uint64 [F,3]."""
import ctypes as C
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'libmc_ref.so')


def available():
    return os.path.exists(LIB)


def marching_cubes(volume, isovalue, truncation):
    lib = C.CDLL(LIB)
    vol = np.ascontiguousarray(volume, dtype=np.float64)
    assert vol.ndim == 3
    lib.mc_ref_run.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_double, C.c_double, C.POINTER(C.c_void_p),
                               C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.mc_ref_free.argtypes = [C.c_void_p]
    pv, pf, nv, nf = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
    rc = lib.mc_ref_run(vol.ctypes.data, vol.shape[0], vol.shape[1], vol.shape[2], float(isovalue), float(truncation),
                        C.byref(pv), C.byref(nv), C.byref(pf), C.byref(nf))
    assert rc == 0
    verts = np.ctypeslib.as_array(C.cast(pv, C.POINTER(C.c_double)), shape=(max(nv.value, 1),))[:nv.value].copy().reshape(-1, 3)
    faces = np.ctypeslib.as_array(C.cast(pf, C.POINTER(C.c_ulong)), shape=(max(nf.value, 1),))[:nf.value].copy().reshape(-1, 3)
    lib.mc_ref_free(pv)
    lib.mc_ref_free(pf)
    return verts, faces.astype(np.uint64)
