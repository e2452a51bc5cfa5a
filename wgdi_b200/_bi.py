"""ctypes binding of libicl_bi_b200.so (include/icl_blockinfo.h): the device part of `-bi`.
No CPU fallback: the call fails when no CUDA device is usable."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libicl_bi_b200.so")
EXPORTS = ("icl_bi_block_stats", "icl_bi_backend")
_LIB = None


class icl_bi_input(C.Structure):
    _fields_ = [("n_genes1", C.c_int64), ("hit_off", C.c_void_p), ("hit_g2", C.c_void_p), ("hit_rank", C.c_void_p),
                ("hit_order2", C.c_void_p), ("n_blocks", C.c_int64), ("blk_off", C.c_void_p), ("mem_g1", C.c_void_p),
                ("mem_g2", C.c_void_p), ("mem_order2", C.c_void_p), ("mem_g2_sorted", C.c_void_p)]


class BiError(RuntimeError):
    pass


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build_bi()
        lib = C.CDLL(LIB_PATH)
        lib.icl_bi_backend.restype = C.c_char_p
        lib.icl_bi_block_stats.argtypes = [C.c_int, C.POINTER(icl_bi_input), C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        if lib.icl_bi_backend() != b"cuda":
            raise BiError(f"{LIB_PATH} is not the CUDA build")
        _LIB = lib
    return _LIB


def block_stats(device, hit_off, hit_g2, hit_rank, hit_order2, blk_off, mem_g1, mem_g2, mem_order2, repeat_number):
    """Per-block integer statistics (see icl_bi_block_stats): homo_sum[B, 5], homo_cnt[B], tandem_rows[B], tandem_uniq[B]."""
    lib = load()
    hit_off = np.ascontiguousarray(hit_off, np.int64)
    blk_off = np.ascontiguousarray(blk_off, np.int64)
    a = [np.ascontiguousarray(x, np.int32) for x in (hit_g2, hit_rank, hit_order2, mem_g1, mem_g2, mem_order2)]
    B = len(blk_off) - 1
    # per block, the subject genes in ascending order (membership test of tandem_ratio's second isin)
    blk_of = np.repeat(np.arange(B, dtype=np.int64), np.diff(blk_off))
    g2s = np.ascontiguousarray(a[4][np.lexsort((a[4], blk_of))], np.int32)
    inp = icl_bi_input(len(hit_off) - 1, hit_off.ctypes.data, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data,
                       B, blk_off.ctypes.data, a[3].ctypes.data, a[4].ctypes.data, a[5].ctypes.data, g2s.ctypes.data)
    homo_sum = np.zeros((B, 5), np.int32)
    homo_cnt = np.zeros(B, np.int32)
    rows = np.zeros(B, np.int64)
    uniq = np.zeros(B, np.int32)
    err = C.create_string_buffer(512)
    rc = lib.icl_bi_block_stats(int(device), C.byref(inp), int(repeat_number), homo_sum.ctypes.data, homo_cnt.ctypes.data,
                                rows.ctypes.data, uniq.ctypes.data, err, 512)
    if rc != 0:
        raise BiError(f"icl_bi error {rc}: {err.value.decode(errors='replace')}")
    return homo_sum, homo_cnt, rows, uniq
