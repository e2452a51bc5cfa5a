"""Typed options and the pandas sort rule of the `-icl` path — pure numpy, no native library behind it, so that the host
ingest, the synthetic-data generator and the benchmark's CPU arm can use them without loading anything CUDA."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class ChainParams:
    """Typed `[collinearity]` options of the DP engine (wgdi/collinearity.py:6-28)."""
    mg1: int = 40
    mg2: int = 40
    grading: tuple = (50, 40, 25)
    pvalue: float = 1.0
    coverage_ratio: float = 0.8
    over_gap: int = 3            # inner over_gap (collinearity.py:13); not a usable conf key (string bug)

    @classmethod
    def from_options(cls, options):
        o = {str(k): v for k, v in options}
        p = cls()
        if "grading" in o:
            p.grading = tuple(int(k) for k in str(o["grading"]).split(","))
        if "mg" in o:
            p.mg1, p.mg2 = [int(k) for k in str(o["mg"]).split(",")]
        p.pvalue = float(o.get("pvalue", 1))
        p.coverage_ratio = float(o.get("coverage_ratio", 0.8))
        return p


def nargsort_desc(a: np.ndarray) -> np.ndarray:
    """pandas.core.sorting.nargsort(a, kind='quicksort', ascending=False) for NaN-free `a`:
    reverse, numpy's unstable argsort, map back, reverse (collinearity.py:49,60; run_colliearity.py:67)."""
    n = len(a)
    r = a[::-1].argsort(kind="quicksort")
    return (n - 1 - r)[::-1]


def compact_points(pair_id, loc1, loc2, grading, bounds=None):
    """The four point arrays in the narrowest unsigned dtype that holds them (uint8 / uint16, else int32): what the ingest
    hands to `IclEngine.upload`, so that a point crosses PCIe in 7 bytes instead of 16 where gene orders and pair ids fit
    16 bits and gradings 8 (include/icl_b200.h, icl_upload_typed).  Negative values stay int32 and are rejected by the upload.
    bounds: optional (lowest, highest) value per array that the caller knows to hold (the ingest does: pair ids are
    0..pairs-1, positions are gene orders, gradings the conf's values) — saves the two passes that look for them."""
    def narrow(x, b):
        x = np.asarray(x)
        if len(x) == 0:
            return np.ascontiguousarray(x, np.int32)
        lo, hi = b if b is not None else (x.min(), x.max())
        if lo < 0:
            return np.ascontiguousarray(x, np.int32)
        hi = int(hi)
        return np.ascontiguousarray(x, np.uint8 if hi < 256 else np.uint16 if hi < 65536 else np.int32)
    b = bounds if bounds is not None else (None, None, None, None)
    return narrow(pair_id, b[0]), narrow(loc1, b[1]), narrow(loc2, b[2]), narrow(grading, b[3])
