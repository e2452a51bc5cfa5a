"""Pipelined lanes: hide the device phases behind the host rank.

Between `icl_dp` and `icl_extract` the host must run numpy's own argsort (the reference's tie order, see
engine.py); on C2 that is about half of a pass and the GPU idles meanwhile.  Chromosome pairs are independent, so
`PipelinedEngine` cuts a comparison into K lanes of consecutive pair ids (balanced by point count), each with its
own context and stream on the same GPU:

    lane 0:  dp ──► rank ──► extract
    lane 1:  dp ─────────► rank ──► extract
    lane 2:  dp ──────────────────► rank ──► extract           (rank phases take turns: each gets all sort threads)

so a lane's extraction runs while the next lane is being ranked.  Lanes hold consecutive pair ranges, hence the
merged result is the plain concatenation of the lanes' blocks — the reference's emission order.
The points are uploaded once (root context) and split on the device (`icl_upload_from`).
"""
from __future__ import annotations

import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .engine import Blocks, ChainParams, IclEngine


class PipelinedEngine:
    def __init__(self, device: int = 0, lanes: int = 4, sort_threads: int | None = None):
        self.root = IclEngine(device, sort_threads)
        self.lanes = [IclEngine(device, sort_threads) for _ in range(max(1, int(lanes)))]
        self.lib = self.root.lib
        self.device = device
        self._threads = ThreadPoolExecutor(len(self.lanes))
        self._rank_turn = threading.Lock()
        self.n = 0
        self.n_pairs = 0
        self.bounds = None
        self.host_ms = {}

    # the benchmark allocates its pinned input buffers through the engine
    def _buf(self, name, n, dtype):
        return self.root._buf(name, n, dtype)

    def close(self):
        for e in self.lanes:
            e.close()
        self.root.close()
        self._threads.shutdown()

    def upload(self, pair_id, loc1, loc2, grading, n_pairs):
        self.root.upload(pair_id, loc1, loc2, grading, n_pairs)
        self.n, self.n_pairs = self.root.n, int(n_pairs)
        counts = np.zeros(max(self.n_pairs, 1), np.int64)
        self.root._check(self.lib.icl_pair_counts(self.root.ctx, counts.ctypes.data))
        cum = np.r_[0, np.cumsum(counts[:self.n_pairs])]
        # consecutive pair ranges with about the same number of points each
        K = len(self.lanes)
        targets = np.arange(1, K) * cum[-1] / K
        cuts = np.searchsorted(cum, targets, side="left")
        self.bounds = np.maximum.accumulate(np.r_[0, np.clip(cuts, 0, self.n_pairs), self.n_pairs]).astype(np.int64)

        def fill(i):
            e = self.lanes[i]
            lo, hi = int(self.bounds[i]), int(self.bounds[i + 1])
            e._check(self.lib.icl_upload_from(e.ctx, self.root.ctx, lo, hi))
            e.n_pairs = hi - lo
            e.n = int(cum[hi] - cum[lo])
        list(self._threads.map(fill, range(K)))

    def run_resident(self, params: ChainParams, min_members: int = 1, copy: bool = True) -> Blocks:
        t0 = time.perf_counter()
        rank_ms = [0.0] * len(self.lanes)

        def lane(i):
            e = self.lanes[i]
            if e.n == 0:
                return None
            kf, kr = e.dp(params)
            with self._rank_turn:                        # one lane ranks at a time, with all sort threads
                t = time.perf_counter()
                af, ar = e.rank(kf, kr)
                rank_ms[i] = (time.perf_counter() - t) * 1e3
            return e.extract(af, ar, min_members, copy=False)
        parts = list(self._threads.map(lane, range(len(self.lanes))))
        t1 = time.perf_counter()
        blocks = self._concat(parts, copy)
        self.host_ms = dict(rank=sum(rank_ms), merge=(time.perf_counter() - t1) * 1e3, total=(time.perf_counter() - t0) * 1e3)
        return blocks

    def chain(self, pair_id, loc1, loc2, grading, n_pairs, params: ChainParams, min_members: int = 1,
              copy: bool = True) -> Blocks:
        self.upload(pair_id, loc1, loc2, grading, n_pairs)
        return self.run_resident(params, min_members, copy)

    def _concat(self, parts, copy):
        """Lanes hold consecutive pair ranges: concatenation is the emission order."""
        live = [(i, p) for i, p in enumerate(parts) if p is not None and len(p)]
        if not live:
            z32, z64 = np.zeros(0, np.int32), np.zeros(0, np.int64)
            return Blocks(z32, z32, z64, z32, z32, z32, z64, z32, z32, np.zeros(0, np.float64), np.zeros(1, np.int64), z32, z32, z32)
        if len(live) == 1 and live[0][0] == 0 and not copy:
            return live[0][1]
        cat = lambda f: np.concatenate([getattr(p, f) for _, p in live])
        pair = np.concatenate([p.pair + np.int32(self.bounds[i]) for i, p in live])
        sizes = [len(p.loc1) for _, p in live]
        base = np.r_[0, np.cumsum(sizes)][:-1]
        moff = np.concatenate([p.member_off[:-1] + b for (_, p), b in zip(live, base)] + [np.array([sum(sizes)], np.int64)])
        return Blocks(pair, cat("dir"), cat("score"), cat("m"), cat("L"), cat("N"), cat("N1"), cat("L1"), cat("L2"),
                      cat("p_raw"), moff, cat("loc1"), cat("loc2"), cat("member_idx"))

    def timings(self):
        """Device phase times summed over lanes (the lanes overlap, so the sum exceeds the wall time)."""
        out = {}
        for e in self.lanes:
            if e.n:
                for k, v in e.timings().items():
                    out[k] = out.get(k, 0.0) + v
        return out

    def counters(self):
        c = [e.counters() for e in self.lanes if e.n]
        if not c:
            return self.root.counters()
        return dict(edges=sum(x["edges"] for x in c), launches_dp=sum(x["launches_dp"] for x in c),
                    launches_extract=sum(x["launches_extract"] for x in c), rows=sum(x["rows"] for x in c),
                    max_span=max(x["max_span"] for x in c), chain_kernel=c[0]["chain_kernel"])
