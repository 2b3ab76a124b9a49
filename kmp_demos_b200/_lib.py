"""ctypes binding of libkmpx.so (include/kmpx.h).  There is no CPU fallback: if the library has not been
built, or no CUDA device is present, every numerical entry point of the package raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libkmpx.so')

LAYOUT_POINT = 0
LAYOUT_COMPONENT = 1
UPLO_FULL = 0
UPLO_LOWER = 1
SMALL_MAX = 512

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_d = C.c_double

# name -> argtypes; every function returns int unless listed in _RESTYPES
_SIGNATURES = {
    'kmpx_version': [],
    'kmpx_last_error': [],
    'kmpx_launch_count': [],
    'kmpx_debug_phases': [_p, _i],
    'kmpx_debug_predict_stats': [_p, _i],
    'kmpx_debug_predict_schedule': [_i, _i, _i, _p, _i],
    'kmpx_predict_skip_threshold': [_d],
    'kmpx_predict_tile_queries': [_i],
    'kmpx_debug_set': [_i],
    'kmpx_debug_exp': [_p, _p, _ll, _i, _p],
    'kmpx_kernel_build': [_p, _p, _p, _p, _i, _i, _i, _i, _i, _ll, _d, _d, _i, _i, _i, _p],
    'kmpx_factor_struct_applicable': [_i, _i, _i],
    'kmpx_factor_struct': [_p, _p, _p, _p, _i, _ll, _i, _i, _i, _i, _d, _d, _p, _p],
    'kmpx_potrf_workspace_bytes': [_i, _i],
    'kmpx_potrf': [_p, _i, _ll, _p, _i, _i, _p, _p, _ll, _p],
    'kmpx_trtri_workspace_bytes': [_i, _i],
    'kmpx_trtri': [_p, _i, _ll, _p, _i, _i, _p, _ll, _p],
    'kmpx_solve_alpha': [_p, _i, _ll, _p, _i, _i, _i, _i, _p, _p, _i, _p, _p],
    'kmpx_inverse_workspace_bytes': [_i, _i],
    'kmpx_inverse': [_p, _i, _ll, _p, _i, _i, _p, _p, _ll, _p],
    'kmpx_apply_inverse': [_p, _i, _ll, _p, _i, _i, _i, _i, _p, _p, _i, _p, _p],
    'kmpx_sys_size': [_i, _i, _i],
    'kmpx_sys_sizes': [_p, _i, _i, _i, _p, _p],
    'kmpx_kmp_predict': [_p, _i, _ll, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p, _ll, _i, _d, _d, _i, _p, _p, _p],
    'kmpx_kmp_predict_persistent': [_p, _i, _ll, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p, _ll, _i, _d, _d, _i, _p, _p, _p],
    'kmpx_inc_workspace_doubles': [_i, _i, _i],
    'kmpx_inc_base': [_p, _i, _ll, _p, _p, _i, _i, _i, _i, _p, _i, _d, _p, _p, _p],
    'kmpx_inc_adapt': [_p, _p, _i, _ll, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _i, _d, _d, _d, _d, _i, _p, _i,
                       _p, _ll, _p, _p, _p],
    'kmpx_inc_grouped_workspace_doubles': [_i, _i, _i],
    'kmpx_inc_adapt_grouped': [_p, _p, _i, _ll, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _i, _p, _p, _p, _i, _d, _d, _d, _d, _i, _p, _i,
                               _p, _ll, _p, _p, _p],
    'kmpx_kmp_predict_gemm_workspace_bytes': [_i, _i, _i],
    'kmpx_kmp_predict_gemm': [_p, _i, _p, _p, _i, _i, _i, _p, _i, _d, _d, _p, _ll, _p, _p, _p],
    'kmpx_kmp_predict_mean': [_p, _p, _p, _i, _i, _i, _i, _p, _ll, _i, _d, _i, _p, _p],
    'kmpx_waypoint_insert': [_p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _d, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    'kmpx_gmr': [_p, _p, _p, _i, _i, _i, _p, _i, _d, _p, _p, _p],
    'kmpx_gmm_workspace_bytes': [_i, _i, _i],
    'kmpx_gmm_em_pass': [_p, _ll, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _p],
    'kmpx_gmm_label_stats': [_p, _ll, _i, _i, _p, _p, _p, _p, _ll, _p],
    'kmpx_gmm_label_sums': [_p, _ll, _i, _i, _p, _p, _p, _p, _ll, _p],
    'kmpx_gmm_finalize': [_p, _d, _i, _i, _i, _d, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    'kmpx_gmm_finalize_check': [_p, _d, _i, _i, _d, _p, _p, _p, _p, _p, _p, _p, _p, _d, _i, _p, _p],
    'kmpx_rows_center_sqnorm': [_p, _ll, _i, _p, _p, _p],
    'kmpx_kpp_candidates': [_p, _p, _ll, _i, _p, _i, _p, _p, _p, _p, _ll, _p],
    'kmpx_kmeans_assign': [_p, _ll, _i, _i, _p, _p, _p, _p, _p],
    'kmpx_kpp_sample_workspace_bytes': [_ll],
    'kmpx_kpp_sample': [_p, _ll, _p, _i, _p, _p, _ll, _ll, _i, _p, _p, _p, _ll, _p],
    'kmpx_kpp_gather_rows': [_p, _i, _p, _i, _ll, _p, _p],
    'kmpx_kpp_candidates_rows': [_p, _p, _ll, _i, _p, _i, _p, _p, _p, _p, _ll, _p],
    'kmpx_kpp_choose': [_p, _i, _i, _i, _p, _i, _i, _ll, _i, _p, _ll, _p, _p, _p, _p, _p, _p, _p, _p],
    'kmpx_lloyd_update': [_p, _i, _i, _p, _p, _p, _p, _i, _p, _p],
    'kmpx_refkmeans_assign': [_p, _ll, _i, _i, _p, _p, _p],
    'kmpx_refkmeans_update': [_p, _ll, _i, _i, _p, _p, _p],
    'kmpx_quat_normalize': [_p, _p, _ll, _p],
    'kmpx_quat_exp': [_p, _p, _ll, _p],
    'kmpx_quat_log': [_p, _p, _ll, _p],
    'kmpx_quat_mul': [_p, _p, _i, _i, _p, _ll, _p],
    'kmpx_quat_from_rotvec': [_p, _p, _ll, _p],
    'kmpx_factor_wpp_applicable': [_p, _i, _ll, _i, _i],
    'kmpx_potrf_wpp': [_p, _i, _ll, _p, _i, _i, _p, _i, _p],
    'kmpx_trtri_wpp': [_p, _i, _ll, _p, _i, _i, _i, _p],
    'kmpx_trtri_wpp_solve': [_p, _i, _ll, _p, _i, _i, _i, _p, _p, _i, _i, _p, _i, _p],
    'kmpx_dgemm': [_i, _i, _i, _i, _i, _d, _p, _i, _ll, _p, _i, _ll, _d, _p, _i, _ll, _i, _i, _p],
}
_RESTYPES = {
    'kmpx_last_error': C.c_char_p,
    'kmpx_launch_count': _ll,
    'kmpx_potrf_workspace_bytes': _ll,
    'kmpx_trtri_workspace_bytes': _ll,
    'kmpx_inverse_workspace_bytes': _ll,
    'kmpx_gmm_workspace_bytes': _ll,
    'kmpx_kpp_sample_workspace_bytes': _ll,
    'kmpx_inc_workspace_doubles': _ll,
    'kmpx_inc_grouped_workspace_doubles': _ll,
    'kmpx_kmp_predict_gemm_workspace_bytes': _ll,
}
EXPORTED = tuple(_SIGNATURES)

_lib = None


class KmpxError(RuntimeError):
    pass


def load():
    """Loads libkmpx.so (once).  Raises if it was not built - there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing: build it with `python -m kmp_demos_b200.csrc.build` '
                          '(nvcc, sm_100a).  kmp_demos_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _i)
    _lib = lib
    return lib


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('kmp_demos_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
    return torch


def call(name, *args):
    """Calls an int-returning entry point and raises KmpxError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise KmpxError(f'{name} failed with status {rc}: {lib.kmpx_last_error().decode()}')


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(load().kmpx_launch_count())
