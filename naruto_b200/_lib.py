"""ctypes binding of libnaruto_b200.so -- the C-ABI declared in include/naruto_b200.h.

No torch types cross this boundary: tensors are passed as raw device pointers (`tensor.data_ptr()`) and the
stream as `torch.cuda.current_stream().cuda_stream`.  There is no CPU fallback: if the shared library cannot
be loaded the import raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libnaruto_b200.so')

NRT_ABI_VERSION = 4
N_LOSS = 8
N_STATS = 16
N_STATS_SUM = 11
STAT_UNCERT_MIN = 11
LOSS_RGB, LOSS_DEPTH, LOSS_SDF, LOSS_FS, LOSS_UNCERT, LOSS_PSNR, LOSS_UNCERT_MIN = range(7)

c_f = C.c_float
c_fp = C.c_void_p          # device pointers travel as integers


class NrtConfig(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('n_levels', C.c_int32), ('n_features', C.c_int32),
                ('log2_hashmap_size', C.c_int32), ('base_resolution', C.c_int32), ('per_level_scale', C.c_double),
                ('n_bins', C.c_int32), ('hidden_dim', C.c_int32), ('geo_feat_dim', C.c_int32),
                ('hidden_dim_color', C.c_int32), ('bound_min', c_f * 3), ('bound_max', c_f * 3),
                ('uncert_dims', C.c_int32 * 3), ('trunc', c_f), ('sc_factor', c_f), ('near_z', c_f), ('far_z', c_f),
                ('depth_trunc', c_f), ('n_samples_d', C.c_int32), ('n_range_d', C.c_int32), ('range_d', c_f)]


class NrtParams(C.Structure):
    _fields_ = [(n, c_fp) for n in ('grid', 'w1', 'w2', 'w3', 'w4', 'uncert')]


class NrtGrads(C.Structure):
    _fields_ = [(n, c_fp) for n in ('grid', 'w1', 'w2', 'w3', 'w4', 'uncert')]


class NrtRenderOut(C.Structure):
    _fields_ = [(n, c_fp) for n in ('rgb', 'depth', 'depth_var', 'acc', 'disp', 'uncert', 'z_vals', 'raw', 'weights', 'feat', 'masks')]


class NrtPeerTable(C.Structure):
    _fields_ = [('world', C.c_int32), ('rank', C.c_int32), ('bucket', C.c_void_p * 8), ('theta', C.c_void_p * 8),
                ('stats_pad', C.c_void_p * 8), ('flags', C.c_void_p * 8), ('bucket_mc', C.c_void_p), ('theta_mc', C.c_void_p)]


class NrtAdamGroup(C.Structure):
    _fields_ = [('begin', C.c_int64), ('end', C.c_int64), ('lr', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float),
                ('eps', C.c_float), ('weight_decay', C.c_float), ('step_dev', C.c_void_p), ('enabled', C.c_int32), ('keep_grad', C.c_int32)]


# name -> (restype, argtypes); every symbol include/naruto_b200.h declares
_P = C.c_void_p
SIGNATURES = {
    'nrt_last_error': (C.c_char_p, []),
    'nrt_abi_version': (C.c_int, []),
    'nrt_plan_create': (C.c_int, [C.POINTER(NrtConfig), C.POINTER(_P)]),
    'nrt_plan_destroy': (None, [_P]),
    'nrt_plan_sizes': (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'nrt_plan_levels': (C.c_int, [_P, C.POINTER(c_f), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'nrt_encode_fwd': (C.c_int, [_P, c_fp, c_fp, C.c_int64, c_fp, _P]),
    'nrt_encode_bwd': (C.c_int, [_P, c_fp, c_fp, C.c_int64, c_fp, c_fp, c_fp, _P]),
    'nrt_oneblob_fwd': (C.c_int, [_P, c_fp, C.c_int64, c_fp, _P]),
    'nrt_oneblob_bwd': (C.c_int, [_P, c_fp, C.c_int64, c_fp, c_fp, _P]),
    'nrt_decode_fwd': (C.c_int, [_P, C.POINTER(NrtParams), c_fp, C.c_int64, C.c_int, c_fp, c_fp, c_fp, _P]),
    'nrt_sample_z': (C.c_int, [_P, c_fp, C.c_int64, c_fp, C.c_int, C.c_uint64, c_fp, _P]),
    'nrt_composite_fwd': (C.c_int, [_P, c_fp, c_fp, C.c_int64, C.c_int32, C.POINTER(NrtRenderOut), _P]),
    'nrt_render_fwd': (C.c_int, [_P, C.POINTER(NrtParams), c_fp, c_fp, c_fp, C.c_int64, c_fp, c_fp, C.c_int, C.c_uint64,
                                 C.POINTER(NrtRenderOut), _P]),
    'nrt_render_fwd_stats': (C.c_int, [_P, C.POINTER(NrtParams), c_fp, c_fp, c_fp, c_fp, C.c_int64, c_fp, C.c_int, C.c_uint64,
                                       c_fp, C.POINTER(NrtRenderOut), c_fp, c_fp, _P]),
    'nrt_step_begin': (C.c_int, [c_fp, C.c_int32, C.c_uint64, c_fp, _P]),
    'nrt_iteration_begin': (C.c_int, [c_fp, c_fp, C.c_uint64, c_fp, _P]),
    'nrt_mc_workspace_bytes': (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    'nrt_mc_extract': (C.c_int, [c_fp, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, c_fp, _P, C.POINTER(C.c_void_p)]),
    'nrt_mc_result_sizes': (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'nrt_mc_result_copy': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'nrt_mc_result_free': (None, [C.c_void_p]),
    'nrt_goal_aggregate': (C.c_int, [c_fp, c_fp, C.POINTER(C.c_int32), c_fp, C.c_int64, c_fp, C.c_int32, C.c_float, C.c_float,
                                     C.c_float, c_fp, c_fp, c_fp, _P]),
    'nrt_erp_depth2dist': (C.c_int, [c_fp, C.c_int32, C.c_int32, c_fp, c_fp, c_fp, C.c_int32, c_fp, _P]),
    'nrt_erp_depth2dist_analytic': (C.c_int, [c_fp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_float, c_fp, _P]),
    'nrt_debug_read': (C.c_int, [C.c_void_p, C.c_int32]),
    'nrt_debug_stamp': (C.c_int, [c_fp, _P]),
    'nrt_stats_exchange': (C.c_int, [C.POINTER(NrtPeerTable), c_fp, c_fp, c_fp, _P]),
    'nrt_adam_step_groups': (C.c_int, [c_fp, c_fp, c_fp, c_fp, C.POINTER(NrtAdamGroup), C.c_int32, C.c_int, _P]),
    'nrt_adam_step_peers': (C.c_int, [C.POINTER(NrtPeerTable), c_fp, c_fp, C.POINTER(NrtAdamGroup), C.c_int32, C.c_int64, c_fp, c_fp, c_fp,
                                     _P]),
    'nrt_loss_stats_bytes': (C.c_int64, []),
    'nrt_loss_partial': (C.c_int, [_P, C.POINTER(NrtRenderOut), c_fp, c_fp, C.c_int64, c_fp, _P]),
    'nrt_loss_finalize': (C.c_int, [_P, c_fp, c_fp, _P]),
    'nrt_loss_fwd': (C.c_int, [_P, C.POINTER(NrtRenderOut), c_fp, c_fp, C.c_int64, c_fp, c_fp, _P]),
    'nrt_decode_bwd': (C.c_int, [_P, C.POINTER(NrtParams), c_fp, C.c_int64, c_fp, C.POINTER(NrtGrads), c_fp, _P]),
    'nrt_render_bwd_workspace': (C.c_int64, [_P, C.c_int64]),
    'nrt_render_bwd': (C.c_int, [_P, C.POINTER(NrtParams), c_fp, c_fp, c_fp, c_fp, C.c_int64, C.POINTER(NrtRenderOut),
                                 c_fp, c_fp, C.POINTER(NrtGrads), c_fp, _P]),
    'nrt_smooth_workspace': (C.c_int64, [_P, C.c_int32]),
    'nrt_smooth_fwd_bwd': (C.c_int, [_P, c_fp, c_fp, C.c_int32, C.c_double, C.c_double, c_f, c_fp, c_fp, c_fp, C.c_int32, C.c_int32,
                                     _P]),
    'nrt_adam_step': (C.c_int, [c_fp, c_fp, c_fp, c_fp, C.c_int64, C.c_int32, c_fp, c_f, c_f, c_f, c_f, c_f, C.c_int, _P]),
    'nrt_counter_add': (C.c_int, [c_fp, C.c_int32, _P]),
    'nrt_map_volumes': (C.c_int, [_P, C.POINTER(NrtParams), C.POINTER(C.c_int32), c_fp, c_fp, _P]),
    'nrt_camera_rays': (C.c_int, [C.c_int32, C.c_int32, c_f, c_f, c_f, c_f, c_fp, _P]),
    'nrt_pack_frame': (C.c_int, [c_fp, c_fp, c_fp, C.c_int64, c_fp, _P]),
    'nrt_valid_depth_count': (C.c_int, [c_fp, C.c_int64, c_f, c_fp, _P]),
    'nrt_kf_store': (C.c_int, [c_fp, c_fp, C.c_int64, C.c_int32, c_fp, c_fp, _P]),
    'nrt_sample_indices': (C.c_int, [C.c_int64, c_fp, C.c_int64, C.c_uint64, c_fp, _P]),
    'nrt_assemble_rays': (C.c_int, [c_fp, c_fp, C.c_int32, C.c_int32, c_fp, C.c_int64, c_fp, c_fp, C.c_int64, c_fp, C.c_int32,
                                    c_fp, c_fp, c_fp, c_fp, _P]),
    'nrt_active_select_workspace': (C.c_int64, [C.c_int64]),
    'nrt_active_select': (C.c_int, [c_fp, c_fp, c_fp, c_fp, C.c_int64, C.c_int64, c_fp, C.POINTER(C.c_int32), C.POINTER(c_f),
                                    C.c_int32, C.c_int32, C.c_int32, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, _P]),
    'nrt_selftest_umma_raw': (C.c_int, [c_fp, C.c_int32, c_fp, C.c_int32] + [C.c_int32] * 10 + [c_fp, _P]),
    'nrt_selftest_umma': (C.c_int, [C.c_int, c_fp, c_fp, C.c_int32, C.c_int32, C.c_int, c_fp, _P]),
}

_lib = None


class NrtError(RuntimeError):
    pass


def load():
    """Load (building first if the source tree is newer and nvcc is present) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError => the .so is stale / symbol missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.nrt_abi_version() != NRT_ABI_VERSION:
        raise NrtError('libnaruto_b200.so ABI version mismatch; rebuild with python -m naruto_b200.build --force')
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise NrtError(f'naruto_b200 error {rc}: {load().nrt_last_error().decode()}')


def ptr(t):
    """Device pointer of a contiguous fp32/fp64/int32 torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), 'naruto_b200 needs contiguous tensors'
    return t.data_ptr()
