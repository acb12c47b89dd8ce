"""ctypes binding of libspindyn_cuda.so (include/spindyn.h).

This is the Python twin of the Julia `ccall` layer shown in INTEGRATION.md: the
reference is Julia and `julia` is not installed in this image, so the
reference's operator/solver interface is mirrored in Python on top of the same
C ABI.  There is no CPU fallback: if the shared library is missing or no CUDA
device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libspindyn_cuda.so")

SD_F64, SD_C128 = 0, 1
SD_PATH_GENERIC, SD_PATH_TILED, SD_PATH_BLOCK = 0, 1, 2
SD_ERR_ARG, SD_ERR_CUDA, SD_ERR_NOMEM, SD_ERR_NCCL, SD_ERR_ZERO_NORM, SD_ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6


class SdBond(ctypes.Structure):
    """Tuple{Int,Int,Float64} (SpinModel.jl:6-15): 24-byte record, 1-based sites."""
    _fields_ = [("i", ctypes.c_int64), ("j", ctypes.c_int64), ("J", ctypes.c_double)]


class SdComplex(ctypes.Structure):
    _fields_ = [("re", ctypes.c_double), ("im", ctypes.c_double)]


class SpinDynError(RuntimeError):
    """Non-argument failure reported by libspindyn_cuda (CUDA, NCCL, memory)."""


class ZeroNormError(RuntimeError):
    """error("starting vector has zero norm")  (Lanczos.jl:210-212)."""


_vp = ctypes.c_void_p
_i = ctypes.c_int
_u64 = ctypes.c_uint64
_d = ctypes.c_double
_P = ctypes.POINTER

# name -> argtypes; every function returns int except the two noted below.
SIGNATURES = {
    "sd_device_count": [_P(_i)],
    "sd_ctx_create": [_i, _P(_vp)],
    "sd_nccl_unique_id": [_vp],
    "sd_ctx_create_rank": [_i, _i, _i, _vp, _P(_vp)],
    "sd_ctx_free": [_vp],
    "sd_ctx_sync": [_vp],
    "sd_ctx_collect": [_vp],
    "sd_ctx_rank": [_vp, _P(_i), _P(_i)],
    "sd_timer_start": [_vp],
    "sd_timer_stop": [_vp, _P(ctypes.c_float)],
    "sd_launch_count": [_vp, _P(_u64)],
    "sd_model_create": [_vp, _i, _i, _vp, _i, _vp, _i, _vp, _P(_vp)],
    "sd_model_free": [_vp],
    "sd_model_dim": [_vp, _P(_u64)],
    "sd_model_local_range": [_vp, _P(_u64), _P(_u64)],
    "sd_model_shard_bounds": [_vp, _i, _vp],
    "sd_model_info": [_vp, _P(_i), _P(_i), _P(_i)],
    "sd_model_set_path": [_vp, _i],
    "sd_unrank": [_vp, _u64, _u64, _vp],
    "sd_rank": [_vp, _vp, _u64, _vp],
    "sd_vec_alloc": [_vp, _i, _P(_vp)],
    "sd_vec_free": [_vp],
    "sd_vec_dtype": [_vp, _P(_i)],
    "sd_vec_local_len": [_vp, _P(_u64)],
    "sd_vec_upload": [_vp, _vp],
    "sd_vec_download": [_vp, _vp],
    "sd_vec_upload_async": [_vp, _vp],
    "sd_vec_download_async": [_vp, _vp],
    "sd_vec_zero": [_vp],
    "sd_vec_set_onehot": [_vp, _u64],
    "sd_vec_get": [_vp, _vp, _u64, _vp, _vp],
    "sd_vec_fill_seeded": [_vp, _u64, _d],
    "sd_vec_copy": [_vp, _vp],
    "sd_vec_convert": [_vp, _vp],
    "sd_vec_scale": [_vp, SdComplex],
    "sd_vec_axpy": [_vp, SdComplex, _vp],
    "sd_vec_dot": [_vp, _vp, _P(SdComplex)],
    "sd_vec_dotu": [_vp, _vp, _P(SdComplex)],
    "sd_vec_norm": [_vp, _P(_d)],
    "sd_host_alloc": [_P(_vp), _u64],
    "sd_host_free": [_vp],
    "sd_apply_H": [_vp, _vp, _vp],
    "sd_apply_H_dot": [_vp, _vp, _vp, _P(SdComplex)],
    "sd_apply_rescaled_H": [_vp, _vp, _vp, _d, _d],
    "sd_cheb_step": [_vp, _vp, _vp, _vp, _d, _d, _vp, _P(_d), _P(_d), _vp, SdComplex],
    "sd_szq": [_vp, _vp, _vp, _d, _P(_d)],
    "sd_apply_sz_weights": [_vp, _vp, _vp, _vp, _P(ctypes.c_double)],
    "sd_apply_H_host": [_vp, _i, _vp, _vp],
    "sd_vec_observables": [_vp, _vp, _vp],
    "sd_vecset_free": [_vp],
    "sd_vecset_size": [_vp, _P(_i)],
    "sd_vecset_get": [_vp, _i, _P(_vp)],
    "sd_lincomb": [_vp, _vp, _i, _vp, _P(_d)],
    "sd_lanczos_extremal": [_vp, _vp, _i, _d, _i, _vp, _vp, _P(_i)],
    "sd_lanczos_groundstate": [_vp, _vp, _i, _d, _d, _vp, _vp, _P(_i), _P(_vp)],
    "sd_lanczos_tridiag": [_vp, _vp, _i, _d, _vp, _vp, _P(_i), _P(_d)],
    "sd_lanczos_lean": [_vp, _vp, _i, _d, _vp, _vp, _P(_i), _vp, _vp, _P(_d)],
    "sd_kpm_moments": [_vp, _vp, _i, _d, _d, _vp],
    "sd_lanczos_tridiag_szq_batch": [_vp, _vp, _vp, _i, _i, _d, _vp, _vp, _vp, _vp],
    "sd_kpm_moments_szq_batch": [_vp, _vp, _vp, _i, _i, _d, _d, _vp, _vp, _P(_i)],
    "sd_ctx_mem_info": [_vp, _P(_u64), _P(_u64)],
    "sd_krylov_basis": [_vp, _vp, _i, _vp, _vp, _P(_i), _P(_d), _P(_vp)],
    "sd_chebyshev_evolve": [_vp, _vp, _vp, _i, _d, _d, _vp],
}
OTHER_SYMBOLS = ("sd_last_error", "sd_version")

_LIB = None


def lib():
    """Load libspindyn_cuda.so; raises if it was not built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SpinDynError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(libspindyn_cuda has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        L.sd_last_error.restype = ctypes.c_char_p
        L.sd_last_error.argtypes = []
        L.sd_version.restype = ctypes.c_int
        L.sd_version.argtypes = []
        _LIB = L
    return _LIB


def check(rc: int) -> None:
    """Map sd_status to the exception the reference would raise."""
    if rc == 0:
        return
    msg = lib().sd_last_error().decode("utf-8", "replace")
    if rc == SD_ERR_ARG:
        raise ValueError(msg)                    # ArgumentError / DimensionMismatch / AssertionError
    if rc == SD_ERR_ZERO_NORM:
        raise ZeroNormError(msg)
    if rc == SD_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == SD_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise SpinDynError(f"libspindyn_cuda error {rc}: {msg}")
