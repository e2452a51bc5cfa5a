"""ctypes binding of libicl_b200.so — the C ABI declared in include/icl_b200.h.

This is the stub a maintainer of the reference would add to call the CUDA engine from
wgdi/run_colliearity.py.  There is no CPU fallback: when the CUDA
library is missing or no GPU is usable, loading or `icl_create` fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libicl_b200.so")

ICL_T_NAMES = ("sort", "setup", "edges", "chain", "export", "extract", "pvalue", "gather", "keys_d2h")


class IclError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"icl_b200 error {code}: {msg}")
        self.code = code


class icl_params(C.Structure):
    _fields_ = [("mg1", C.c_int32), ("mg2", C.c_int32), ("g0", C.c_int32), ("over_gap", C.c_int32),
                ("coverage_ratio", C.c_double)]


class icl_blocks(C.Structure):
    _fields_ = [("n_blocks", C.c_int64), ("n_members", C.c_int64),
                ("pair", C.POINTER(C.c_int32)), ("dir", C.POINTER(C.c_int32)), ("score", C.POINTER(C.c_int64)),
                ("m", C.POINTER(C.c_int32)), ("L", C.POINTER(C.c_int32)), ("N", C.POINTER(C.c_int32)),
                ("N1", C.POINTER(C.c_int64)), ("L1", C.POINTER(C.c_int32)), ("L2", C.POINTER(C.c_int32)),
                ("p_raw", C.POINTER(C.c_double)), ("member_off", C.POINTER(C.c_int64)),
                ("loc1", C.POINTER(C.c_int32)), ("loc2", C.POINTER(C.c_int32)), ("member_idx", C.POINTER(C.c_int32))]


_LIB = None


def _declare(lib):
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.icl_backend.restype = C.c_char_p
    lib.icl_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.icl_destroy.argtypes = [vp]
    lib.icl_destroy.restype = None
    lib.icl_last_error.argtypes = [vp]
    lib.icl_last_error.restype = C.c_char_p
    lib.icl_host_alloc.argtypes = [vp, i64, C.POINTER(vp)]
    lib.icl_host_free.argtypes = [vp, vp]
    lib.icl_upload.argtypes = [vp, i64, i32, vp, vp, vp, vp]
    lib.icl_upload_typed.argtypes = [vp, i64, i32, vp, i32, vp, i32, vp, i32, vp, i32]
    lib.icl_set_grading_table.argtypes = [vp, i32, vp]
    lib.icl_dp.argtypes = [vp, C.POINTER(icl_params), vp, vp]
    lib.icl_dp_begin.argtypes = [vp, C.POINTER(icl_params), vp, vp, vp]
    lib.icl_dp_end.argtypes = [vp]
    lib.icl_layout.argtypes = [vp, vp, vp]
    lib.icl_rank_device.argtypes = [vp, i32]
    lib.icl_debug_order.argtypes = [vp, vp]
    lib.icl_debug_rank_keys.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.icl_debug_keys.argtypes = [vp, vp]
    lib.icl_extract.argtypes = [vp, vp, vp, i32, C.POINTER(icl_blocks)]
    lib.icl_extract_fields.argtypes = [vp, vp, vp, i32, i32, C.POINTER(icl_blocks)]
    lib.icl_free_blocks.argtypes = [C.POINTER(icl_blocks)]
    lib.icl_free_blocks.restype = None
    lib.icl_debug_dp.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.icl_timings.argtypes = [vp, vp, i32]
    lib.icl_counters.argtypes = [vp, vp, i32]
    lib.icl_cache_stats.argtypes = [vp, i32]
    lib.icl_cache_trim.argtypes = []
    lib.icl_cache_trim.restype = None
    lib.icl_stream.argtypes = [vp]
    lib.icl_stream.restype = vp
    return lib


EXPORTS = ("icl_backend", "icl_create", "icl_destroy", "icl_last_error", "icl_host_alloc", "icl_host_free",
           "icl_upload", "icl_upload_typed", "icl_set_grading_table", "icl_dp", "icl_dp_begin", "icl_dp_end", "icl_layout", "icl_rank_device", "icl_debug_order", "icl_debug_rank_keys", "icl_debug_keys", "icl_extract", "icl_extract_fields", "icl_free_blocks", "icl_debug_dp", "icl_timings",
           "icl_counters", "icl_stream", "icl_cache_stats", "icl_cache_trim")


def load():
    """Load the CUDA library (building it with nvcc first if the .so is absent)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = _declare(C.CDLL(LIB_PATH))
    if lib.icl_backend() != b"cuda":
        raise IclError(-2, f"{LIB_PATH} is not the CUDA build (backend={lib.icl_backend()!r}); refusing to run")
    _LIB = lib
    return lib
