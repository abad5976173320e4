"""ctypes binding of libb200boot.so (C ABI declared in include/b200boot.h).

There is no CPU fallback: if the library is missing or fails to load, every
entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# B200BOOT_LIB_PATH: development aid (tools/try_variants.sh times experimental builds)
LIB_PATH = os.environ.get("B200BOOT_LIB_PATH") or os.path.join(_PKG_DIR, "libb200boot.so")

ABI_VERSION = 1

STAT_MEAN, STAT_VAR, STAT_STD, STAT_RATIO_OF_MEANS, STAT_WEIGHTED_AVERAGE, STAT_PEARSON_R, \
    STAT_MEDIAN, STAT_SUM, STAT_SLOPE, STAT_INTERCEPT, STAT_LINEAR_FIT = range(11)
STAT_OUTPUTS = {STAT_LINEAR_FIT: 2}      # outputs per resample where not 1
CMP_LT, CMP_LE, CMP_GT, CMP_GE, CMP_BETWEEN = range(5)
OP_SUB, OP_ADD, OP_DIV, OP_MUL, OP_LOG_RATIO = range(5)


class Plan(C.Structure):
    _fields_ = [("log2_tile", C.c_int32), ("n_arrays", C.c_int32), ("n_full", C.c_int64),
                ("rem", C.c_int64), ("n_tiles", C.c_int64)]


class RankRule(C.Structure):
    """include/b200boot.h: b200boot_rank_rule."""
    _fields_ = [("k_lo", C.c_int64), ("k_hi", C.c_int64), ("gamma", C.c_double),
                ("loo_k_lo", C.c_int64), ("loo_k_hi", C.c_int64), ("loo_gamma", C.c_double),
                ("mode", C.c_int32), ("reserved", C.c_int32)]


class B200BootError(RuntimeError):
    """A libb200boot call returned a non-zero status."""


_vp, _i64, _u64, _int, _sz = C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_size_t

# name -> (restype, argtypes); mirrors include/b200boot.h one to one
SIGNATURES = {
    "b200boot_version": (_int, []),
    "b200boot_philox_rounds": (_int, []),
    "b200boot_last_error": (C.c_char_p, []),
    "b200boot_make_plan": (_int, [_i64, _int, C.POINTER(Plan)]),
    "b200boot_workspace_bytes": (_sz, [_int, _i64, _int, _i64, _i64, _int]),
    "b200boot_resample_stat": (_int, [_vp, _int, _i64, _int, _u64, _i64, _i64, _vp, _vp, _sz, _vp]),
    "b200boot_gemm_supported": (_int, [_int, _i64]),
    "b200boot_count_matrix": (_int, [_u64, _i64, _i64, _i64, _vp, _vp]),
    "b200boot_counts_from_indices": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "b200boot_finish_columns": (_int, [_int, _i64, _i64, _i64, _vp, _vp, _vp]),
    "b200boot_resample_stat_after": (_int, [_vp, _int, _i64, _int, _u64, _i64, _i64, _vp, _vp, _sz, _vp, _vp]),
    "b200boot_resample_stat_indexed": (_int, [_vp, _int, _i64, _int, _vp, _i64, _vp, _vp, _sz, _vp]),
    "b200boot_columns_block_width": (_int, [_int, _i64]),
    "b200boot_resample_stat_columns": (_int, [_vp, _i64, _i64, _i64, _int, _u64, _i64, _i64, _vp, _vp, _sz, _vp]),
    "b200boot_jackknife_columns": (_int, [_vp, _i64, _i64, _i64, _int, _vp, _vp]),
    "b200boot_resample_median_columns": (_int, [_vp, _i64, _i64, _i64, _u64, _i64, _i64, _vp, _vp, _sz, _vp, _vp, _vp]),
    "b200boot_resample_median_columns_indexed": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _sz, _vp, _vp]),
    "b200boot_resample_median_sorted": (_int, [_vp, _i64, _u64, _i64, _i64, _vp, _vp, _sz, _vp, _vp]),
    "b200boot_jackknife_median_sorted": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "b200boot_fill_indices": (_int, [_u64, _i64, _int, _i64, _i64, _vp, _vp, _sz, _vp]),
    "b200boot_fill_moving_block": (_int, [_u64, _i64, _i64, _int, _i64, _i64, _vp, _vp]),
    "b200boot_resample_stat_moving_block": (_int, [_vp, _int, _i64, _int, _u64, _i64, _int, _i64, _i64, _vp, _vp, _sz, _vp]),
    "b200boot_fill_subsample": (_int, [_u64, _i64, _i64, _i64, _i64, _vp, _vp, _sz, _vp]),
    "b200boot_stream_synchronize": (_int, [_vp]),
    "b200boot_ci_small_supported": (_int, [_int, _int, _i64, _i64, _int]),
    "b200boot_ci_small": (_int, [_vp, _int, _i64, _int, _u64, _i64, _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200boot_rank_columns_ws_bytes": (_sz, [_i64, _i64]),
    "b200boot_rank_columns": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "b200boot_rank_columns_ld": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _sz, _vp]),
    "b200boot_median_rows_indexed_ld": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _vp]),
    "b200boot_median_rows_supported": (_int, [_i64]),
    "b200boot_median_rows_indexed": (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _vp]),
    "b200boot_ci_tail_supported": (_int, [_i64, _int]),
    "b200boot_ci_tail": (_int, [_vp, _i64, _vp, _vp, _int, _int, _vp, _vp, _vp]),
    "b200boot_host_ncdf": (None, [_vp, _vp, _i64]),
    "b200boot_nan_columns": (_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "b200boot_nan_fixup": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp]),
    "b200boot_nan_poison_rows": (_int, [_vp, _i64, _int, _vp, _vp]),
    "b200boot_resample_stat_host": (_int, [_vp, _vp, _int, _i64, _int, _u64, _i64, _i64, _vp, _vp, _sz, _vp, _vp]),
    "b200boot_pval_small": (_int, [_vp, _vp, _vp, _int, _i64, _int, _u64, _i64, _int, C.c_double, C.c_double, _vp, _vp, _vp,
                                   _vp, _sz, _vp]),
    "b200boot_ci_median_small_supported": (_int, [_i64, _i64, _int]),
    "b200boot_ci_median_small": (_int, [_vp, _vp, _vp, _i64, _u64, _i64, _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp,
                                        _sz, _vp]),
    "b200boot_ci_small_host": (_int, [_vp, _vp, _vp, _int, _i64, _int, _u64, _i64, _vp, _int, _int, _int, _vp, _vp, _vp,
                                      _vp, _sz, _vp]),
    "b200boot_tile_counts": (_int, [_u64, _i64, _int, _i64, _i64, _vp, _vp, _vp]),
    "b200boot_tile_counts_packed_ws_bytes": (_sz, [_i64, _int, _i64]),
    "b200boot_tile_counts_packed": (_int, [_u64, _i64, _int, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "b200boot_gather": (_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "b200boot_jackknife": (_int, [_vp, _int, _i64, _int, _vp, _vp, _sz, _vp]),
    "b200boot_jackknife_median_columns": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _sz, _vp, _vp]),
    "b200boot_combine": (_int, [_int, _vp, _vp, _vp, _i64, _vp]),
    "b200boot_jackknife_values": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _sz, _vp]),
    "b200boot_jackknife_values_sorted": (_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "b200boot_jackknife_pooled": (_int, [_int, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "b200boot_abc_identity_averages": (_int, [_vp, _i64, C.c_double, _vp, _vp, _vp]),
    "b200boot_weighted_averages": (_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "b200boot_count": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp]),
    "b200boot_select": (_int, [_vp, _i64, _i64, _vp, _int, _vp, _vp, _sz, _vp]),
    "b200boot_select_begin": (_int, [_vp, _int, _i64, _vp, _sz, _vp]),
    "b200boot_select_hist": (_int, [_vp, _i64, _i64, _int, _int, _vp, _sz, _vp]),
    "b200boot_select_pick": (_int, [_int, _i64, _int, _vp, _sz, _vp]),
    "b200boot_select_end": (_int, [_int, _i64, _vp, _vp, _sz, _vp]),
    "b200boot_select_hist_buffer": (_int, [_int, _i64, _vp, _sz, C.POINTER(_vp), C.POINTER(_i64)]),
    "b200boot_sort_columns": (_int, [_vp, _i64, _i64, _vp, _sz, _vp]),
    "b200boot_count_gemm_ws_bytes": (_sz, [_i64, _i64, _i64]),
    "b200boot_resample_stat_columns_gemm": (_int, [_vp, _i64, _i64, _i64, _int, _u64, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200boot_debug_i8_gemm": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "b200boot_launch_count": (_i64, []),
    "b200boot_profile_enable": (_int, [_int]),
    "b200boot_profile_last_k1_ms": (_int, [C.POINTER(C.c_float)]),
    "b200boot_profile_last_ms": (_int, [C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "b200boot_host_philox4x32": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "b200boot_host_philox4x32_r": (None, [_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "b200boot_host_binomial": (_i64, [_u64, _i64, _i64, _i64, C.c_double]),
    "b200boot_host_binomial_half": (_i64, [_u64, _i64, _i64, _i64]),
    "b200boot_host_tile_counts": (_int, [_u64, _i64, _int, _i64, _vp]),
    "b200boot_host_class_counts": (_int, [_u64, _i64, _i64, _i64, _vp]),
    "b200boot_host_indices": (_int, [_u64, _i64, _int, _i64, _vp]),
}

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Loads (once) and returns the library; raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200BootError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.b200boot_version() != ABI_VERSION:
            raise B200BootError("libb200boot.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = lib().b200boot_last_error()
        raise B200BootError(f"libb200boot status {status}: {msg.decode() if msg else ''}")


def make_plan(n_rows: int, n_arrays: int) -> Plan:
    p = Plan()
    check(lib().b200boot_make_plan(n_rows, n_arrays, C.byref(p)))
    return p
