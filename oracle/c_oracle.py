"""ctypes binding of oracle/libicl_oracle.so (the C restatement in oracle/icl_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .icl_oracle import IclParams, nargsort_desc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libicl_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.icl_oracle_extract.restype = C.c_int64
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def chain(loc1, loc2, grading, mg1, mg2, direction):
    """-> (score, pred, used, depth) as int64/bool arrays, like icl_oracle.chain_dp."""
    loc1, loc2, g = _i32(loc1), _i32(loc2), _i32(grading)
    n = len(loc1)
    score, pred, depth = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32)
    used = np.empty(n, np.uint8)
    lib().icl_oracle_chain(C.c_int64(n), _p(loc1), _p(loc2), _p(g), C.c_int32(mg1), C.c_int32(mg2),
                           C.c_int32(direction), _p(score), _p(pred), _p(depth), _p(used))
    return score.astype(np.int64), pred.astype(np.int64), used.astype(bool), depth.astype(np.int64)


def extract(loc1, loc2, score, pred, used, depth, order, times, params: IclParams):
    """All accepted blocks of one direction: list of dicts (members, a, score, L, N, N1, L1, L2).
    `times` (int32 array) is updated in place."""
    loc1, loc2 = _i32(loc1), _i32(loc2)
    n = len(loc1)
    score32, pred32, depth32 = _i32(score), _i32(pred), _i32(depth)
    used8 = np.ascontiguousarray(used, np.uint8)
    order32 = _i32(order)
    cap = n // 3 + 2
    blk_off = np.zeros(cap + 1, np.int64)
    members = np.zeros(max(n, 1), np.int32)
    top, L, N, L1, L2 = (np.zeros(cap, np.int32) for _ in range(5))
    N1 = np.zeros(cap, np.int64)
    a = np.zeros(cap, np.float64)
    assert times.dtype == np.int32
    nb = lib().icl_oracle_extract(C.c_int64(n), _p(loc1), _p(loc2), _p(score32), _p(pred32), _p(depth32), _p(used8),
                                  _p(order32), _p(times), C.c_int32(params.grading[0]), C.c_int32(params.over_gap),
                                  C.c_double(params.coverage_ratio), _p(blk_off), _p(members),
                                  _p(top), _p(L), _p(N), _p(N1), _p(L1), _p(L2), _p(a))
    out = []
    for b in range(nb):
        out.append(dict(members=members[blk_off[b]:blk_off[b + 1]].astype(np.int64), a=float(a[b]),
                        score=int(score32[top[b]]), top=int(top[b]), L=int(L[b]), N=int(N[b]), N1=int(N1[b]),
                        L1=int(L1[b]), L2=int(L2[b])))
    return out


def collinearity_run(loc1, loc2, grading, options=(), order_fn=nargsort_desc, orders=None):
    """Inner seam with the C kernels: [(loc1[], loc2[], p, score, direction)] like icl_oracle.collinearity_run.
    `orders` = (order_fwd, order_rev) overrides the numpy call (used with golden tie orders)."""
    params = options if isinstance(options, IclParams) else IclParams(options)
    loc1 = np.asarray(loc1, np.int64)
    loc2 = np.asarray(loc2, np.int64)
    times = np.ones(len(loc1), np.int32)
    data = []
    for di, direction in enumerate((+1, -1)):
        score, pred, used, depth = chain(loc1, loc2, grading, params.mg1, params.mg2, direction)
        order = orders[di] if orders is not None else order_fn(score)
        for b in extract(loc1, loc2, score, pred, used, depth, order, times, params):
            p = round(np.float64(b["a"]), 4)
            if p > params.pvalue:
                continue
            mm = b["members"]
            data.append((loc1[mm], loc2[mm], p, b["score"], direction))
    return data


class Batch:
    """All chromosome pairs of one comparison, concatenated, for the multi-threaded CPU baseline."""

    def __init__(self, off, loc1, loc2, grading, params: IclParams, threads=None):
        self.off = np.ascontiguousarray(off, np.int64)
        self.loc1, self.loc2, self.g = _i32(loc1), _i32(loc2), _i32(grading)
        self.params = params
        self.threads = threads or os.cpu_count() or 1
        self.n = len(self.loc1)
        self.n_pairs = len(self.off) - 1
        sizes = np.diff(self.off)
        self.pair_order = _i32(np.argsort(-sizes, kind="stable"))                 # longest first
        self.item_order = _i32(np.repeat(self.pair_order, 2) * 2 + np.tile([0, 1], self.n_pairs))

    def chain(self):
        n = self.n
        self.dp = [dict(score=np.empty(n, np.int32), pred=np.empty(n, np.int32), depth=np.empty(n, np.int32),
                        used=np.empty(n, np.uint8)) for _ in range(2)]
        f, r = self.dp
        lib().icl_oracle_chain_batch(C.c_int32(self.n_pairs), C.c_int32(self.threads), _p(self.item_order),
                                     _p(self.off), _p(self.loc1), _p(self.loc2), _p(self.g),
                                     C.c_int32(self.params.mg1), C.c_int32(self.params.mg2),
                                     _p(f["score"]), _p(f["pred"]), _p(f["depth"]), _p(f["used"]),
                                     _p(r["score"]), _p(r["pred"]), _p(r["depth"]), _p(r["used"]))
        return self.dp

    def rank(self, pool=None):
        """numpy's own unstable argsort per (pair, direction) segment, on int64 like the reference."""
        self.order = [np.empty(self.n, np.int32) for _ in range(2)]

        def one(args):
            d, p = args
            a, b = self.off[p], self.off[p + 1]
            self.order[d][a:b] = nargsort_desc(self.dp[d]["score"][a:b].astype(np.int64))
        items = [(d, p) for p in self.pair_order for d in (0, 1)]
        if pool is None:
            for it in items:
                one(it)
        else:
            list(pool.map(one, items))
        return self.order

    def extract(self):
        n, P = self.n, self.n_pairs
        sizes = np.diff(self.off)
        cap = 2 * (sizes // 3) + 2
        blk_base = np.r_[0, np.cumsum(cap)].astype(np.int64)
        tot = int(blk_base[-1])
        self.blk_base = blk_base
        self.n_blk = [np.zeros(P, np.int64), np.zeros(P, np.int64)]
        self.blk_off = np.zeros(tot + 2 * P + 2, np.int64)
        self.members = np.zeros(2 * n + 1, np.int32)
        self.top, self.L, self.N, self.L1, self.L2 = (np.zeros(tot + 1, np.int32) for _ in range(5))
        self.N1 = np.zeros(tot + 1, np.int64)
        self.a = np.zeros(tot + 1, np.float64)
        f, r = self.dp
        lib().icl_oracle_extract_batch(
            C.c_int32(P), C.c_int32(self.threads), _p(self.pair_order), _p(self.off), _p(self.loc1), _p(self.loc2),
            _p(f["score"]), _p(f["pred"]), _p(f["depth"]), _p(f["used"]),
            _p(r["score"]), _p(r["pred"]), _p(r["depth"]), _p(r["used"]),
            _p(self.order[0]), _p(self.order[1]),
            C.c_int32(self.params.grading[0]), C.c_int32(self.params.over_gap), C.c_double(self.params.coverage_ratio),
            _p(blk_base), _p(self.n_blk[0]), _p(self.n_blk[1]), _p(self.blk_off), _p(self.members),
            _p(self.top), _p(self.L), _p(self.N), _p(self.N1), _p(self.L1), _p(self.L2), _p(self.a))

    def blocks(self, pair):
        """[(members_global_idx, a, score, direction)] of one pair in reference order (all accepted blocks)."""
        out = []
        a0 = int(self.off[pair])
        n = int(self.off[pair + 1] - a0)
        b = int(self.blk_base[pair])
        for d in range(2):
            nb = int(self.n_blk[d][pair])
            offs = self.blk_off[b + 2 * pair + d: b + 2 * pair + d + nb + 1]
            area = 2 * a0 + d * n
            for j in range(nb):
                mm = self.members[area + offs[j]: area + offs[j + 1]].astype(np.int64) + a0
                out.append((mm, float(self.a[b + j]), int(self.dp[d]["score"][a0 + self.top[b + j]]), +1 if d == 0 else -1))
            b += nb
        return out
