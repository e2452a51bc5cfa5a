"""Host driver of the CUDA chaining engine: one `IclEngine` per GPU.

The engine takes the graded points of a whole comparison (all chromosome pairs), runs the
two device phases of include/icl_b200.h and, between them, obtains the extraction order
from numpy exactly the way the reference does through pandas
(`sort_values(by=['score1'], ascending=False)`, wgdi/collinearity.py:49,60 ->
pandas.core.sorting.nargsort -> numpy.argsort(kind='quicksort')).  That sort is unstable
and CPU dependent for tied scores (SURVEY.md hazard H1).  The device holds exact
reproductions of the three algorithms numpy dispatches to on x86 (AVX-512, AVX2, scalar);
the engine checks them against the running numpy once per process and lets the one that
matches rank on the device (`numpy_sort_kind`).  When none matches, the very same numpy
call is made on the host on the int64 score arrays the device returns.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import IclError, icl_blocks, icl_params

FIELDS = {"all": 0, "writer": 1}         # ICL_FIELDS_ALL / ICL_FIELDS_WRITER of include/icl_b200.h
from .params import ChainParams, nargsort_desc  # noqa: F401  (re-exported: the reference-facing names live in params.py)

try:                                    # numpy's own argsort routine, GIL-free (wgdi_b200/csrc/npsort.c)
    from . import _npsort
except ImportError:                     # not built: one numpy call per segment instead (same results)
    _npsort = None
_NPSORT_OK = None
_NP_KIND = {}            # device -> which device reproduction of numpy's argsort matches this process (IclEngine.numpy_sort_kind)


def _npsort_ok():
    """True when the helper reproduces numpy.argsort(kind='quicksort') on tie-heavy int64 arrays of several
    sizes in this process (checked once); any difference disables it."""
    global _NPSORT_OK
    if _NPSORT_OK is None:
        ok = _npsort is not None
        if ok:
            rng = np.random.default_rng(12345)
            sizes = np.array([0, 1, 2, 7, 16, 17, 64, 129, 255, 256, 257, 1000, 5000, 40000], np.int64)
            off = np.r_[0, np.cumsum(sizes)].astype(np.int64)
            keys = rng.integers(25, 60, int(off[-1])).astype(np.int32)
            out = np.empty(len(keys), np.int32)
            try:
                _npsort.seg_argsort(keys, off, np.arange(len(sizes), dtype=np.int32), out, 3)
                ok = all((out[a:b] == keys[a:b].astype(np.int64).argsort(kind="quicksort")).all()
                         for a, b in zip(off[:-1], off[1:]))
            except Exception:
                ok = False
        _NPSORT_OK = ok
    return _NPSORT_OK


@dataclass
class Blocks:
    """All accepted blocks of a comparison in the reference's emission order (see icl_blocks).  With `IclEngine.fields =
    "writer"` only what the collinearity file is written from comes back from the device (pair, score, m, p_raw,
    member_off, loc1, loc2); the other arrays are None."""
    pair: np.ndarray
    dir: np.ndarray
    score: np.ndarray
    m: np.ndarray
    L: np.ndarray
    N: np.ndarray
    N1: np.ndarray
    L1: np.ndarray
    L2: np.ndarray
    p_raw: np.ndarray
    member_off: np.ndarray
    loc1: np.ndarray
    loc2: np.ndarray
    member_idx: np.ndarray

    def __len__(self):
        return len(self.pair)

    def p_value(self):
        """round(a, 4) with numpy's float64 rounding, as collinearity.py:153 does."""
        return np.round(self.p_raw, 4)


class _Pinned:
    """numpy view of page-locked host memory owned by the library."""

    def __init__(self, eng, n, dtype):
        self.eng = eng
        self.ptr = C.c_void_p()
        nbytes = max(int(n), 1) * np.dtype(dtype).itemsize
        eng._check(eng.lib.icl_host_alloc(eng.ctx, nbytes, C.byref(self.ptr)))
        buf = (C.c_char * nbytes).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(n))

    def free(self):
        if self.ptr:
            self.array = None
            self.eng.lib.icl_host_free(self.eng.ctx, self.ptr)
            self.ptr = None


class IclEngine:
    def __init__(self, device: int = 0, sort_threads: int | None = None):
        self.lib = _lib.load()
        self.ctx = C.c_void_p()
        rc = self.lib.icl_create(int(device), C.byref(self.ctx))
        if rc != 0:
            raise IclError(rc, self.lib.icl_last_error(None).decode())
        self.device = device
        self.n = 0
        self.n_pairs = 0
        self._pinned = {}
        self.sort_threads = sort_threads or min(32, os.cpu_count() or 1)
        # After sorting, the helper drops the key buffer's lines from the CPU caches (CLFLUSHOPT) so that the next
        # device-to-host DMA does not have to snoop them out: 15 MB in 0.29 ms instead of 0.84 ms (csrc/npsort.c).
        default_flush = "3" if (_npsort is not None and hasattr(_npsort, "has_clflushopt") and _npsort.has_clflushopt()) else "0"
        self.flush_keys = int(os.environ.get("ICL_B200_FLUSH_KEYS", default_flush))
        self._pool = ThreadPoolExecutor(self.sort_threads) if self.sort_threads > 1 else None
        # overlap the host rank with the DP kernel (icl_dp_begin); ICL_B200_STREAM=0 runs the two one after the other
        self.stream_rank = os.environ.get("ICL_B200_STREAM", "1") != "0"
        # who ranks: "auto" = the device when it provably gives numpy's own permutation in this process (numpy's AVX-512,
        # AVX2 or scalar argsort, checked against the running numpy), else the host; "host" = always numpy on the host;
        # "numpy_scalar" / "numpy_avx512" / "numpy_avx2" = that device reproduction regardless of what numpy does here
        self.rank_mode = os.environ.get("ICL_B200_RANK", "auto")
        self.fields = "all"              # what `extract` brings back: "all" arrays of icl_blocks, or what the file "writer" reads
        self.host_rank_fallbacks = 0        # comparisons the device rank handed back to numpy (std::sort's heap-sort branch)
        # ICL_B200_RANK_VERIFY=K (1 = 32): after every device rank, K segments (the longest four, the rest at random) are
        # re-sorted with the running numpy and compared; a difference sends this comparison — and from then on the process —
        # to the host rank and is counted here
        v = os.environ.get("ICL_B200_RANK_VERIFY", "0")
        self.rank_verify = (32 if v == "1" else max(0, int(v))) if v.isdigit() else 0
        self.rank_verify_failures = 0
        self.rank_verified_segments = 0
        self._verify_rng = np.random.default_rng(977)
        self.host_ms = {}

    # ---- plumbing ------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise IclError(rc, self.lib.icl_last_error(self.ctx).decode())

    def close(self):
        if self.ctx:
            for p in self._pinned.values():
                p.free()
            self._pinned = {}
            self.lib.icl_destroy(self.ctx)
            self.ctx = None
        if self._pool:
            self._pool.shutdown()
            self._pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _buf(self, name, n, dtype):
        p = self._pinned.get(name)
        if p is None or len(p.array) < n:
            if p is not None:
                p.free()
            p = _Pinned(self, int(n * 1.1) + 16, dtype)
            self._pinned[name] = p
        return p.array[:n]

    # ---- device phases -------------------------------------------------------------------------
    def upload(self, pair_id, loc1, loc2, grading, n_pairs):
        """Arrays of uint8 / uint16 cross PCIe as they are (icl_upload_typed; see `compact_points`); anything else is
        passed as int32."""
        a = [np.ascontiguousarray(x) if np.asarray(x).dtype in (np.uint8, np.uint16) else np.ascontiguousarray(x, np.int32)
             for x in (pair_id, loc1, loc2, grading)]
        n = len(a[0])
        if not all(len(x) == n for x in a):
            raise ValueError("pair_id, loc1, loc2, grading must have equal length")
        if all(x.dtype == np.int32 for x in a):
            self._check(self.lib.icl_upload(self.ctx, n, int(n_pairs), *[x.ctypes.data for x in a]))
        else:
            args = []
            for x in a:
                args += [x.ctypes.data, x.dtype.itemsize]
            self._check(self.lib.icl_upload_typed(self.ctx, n, int(n_pairs), *args))
        self.n, self.n_pairs = n, int(n_pairs)
        self.upload_bytes = sum(x.nbytes for x in a)

    def set_grading_table(self, values=None):
        """Sweeps over `grading`: with a table the `grading` array of `upload` holds rank-class ids (0 red, 1 blue, 2 rest,
        wgdi/run_colliearity.py:64-86) and every later DP chains with values[class] — same upload, new gradings.
        None restores plain grading values."""
        if values is None:
            self._check(self.lib.icl_set_grading_table(self.ctx, 0, None))
        else:
            v = np.ascontiguousarray(values, np.int32)
            self._check(self.lib.icl_set_grading_table(self.ctx, len(v), v.ctypes.data))

    def dp(self, params: ChainParams):
        """Both DP directions; returns (key_fwd, key_rev): int32 scores, every pair segment in REVERSED point
        order (widened to int64: the array pandas' nargsort(ascending=False) hands to numpy.argsort)."""
        prm = icl_params(int(params.mg1), int(params.mg2), int(params.grading[0]), int(params.over_gap),
                         float(params.coverage_ratio))
        keys = self._buf("keys", 2 * self.n, np.int32)               # forward | anti-diagonal, one pinned buffer
        sf, sr = keys[:self.n], keys[self.n:]
        self._check(self.lib.icl_dp(self.ctx, C.byref(prm), sf.ctypes.data, sr.ctypes.data))
        self.pair_off = np.zeros(self.n_pairs + 1, np.int64)
        self._check(self.lib.icl_layout(self.ctx, self.pair_off.ctypes.data, None))
        return sf, sr

    def layout(self):
        perm = np.zeros(self.n, np.int32)
        off = np.zeros(self.n_pairs + 1, np.int64)
        self._check(self.lib.icl_layout(self.ctx, off.ctypes.data, perm.ctypes.data))
        return off, perm

    def rank(self, key_fwd, key_rev):
        """numpy.argsort(kind='quicksort') of every pair segment of the arrays `dp` returned — the very sort the
        reference reaches through pandas (collinearity.py:49,60).  Uses numpy's own compiled routine from a
        thread pool without the GIL when the helper module is available (verified against numpy.argsort at
        start-up), otherwise one numpy call per segment."""
        both = self._buf("argsort", 2 * self.n, np.int32)           # forward | anti-diagonal, one pinned buffer
        af, ar = both[:self.n], both[self.n:]
        off = self.pair_off
        sizes = np.diff(off)
        if _npsort_ok():
            # one pass over all 2P segments (longest first) so that the pool drains evenly
            if key_rev.ctypes.data == key_fwd.ctypes.data + 4 * self.n:
                keys2 = np.frombuffer((C.c_int32 * (2 * self.n)).from_address(key_fwd.ctypes.data), np.int32) if self.n else key_fwd
                off2 = np.concatenate([off, off[1:] + self.n])
                order = np.argsort(-np.concatenate([sizes, sizes]), kind="stable").astype(np.int32)
                _npsort.seg_argsort(keys2, off2, order, both, self.sort_threads, self.flush_keys)
                return af, ar
            order = np.argsort(-sizes, kind="stable").astype(np.int32)
            _npsort.seg_argsort(key_fwd, off, order, af, self.sort_threads)
            _npsort.seg_argsort(key_rev, off, order, ar, self.sort_threads)
            return af, ar
        pairs = np.argsort(-sizes, kind="stable")
        pairs = pairs[sizes[pairs] > 0]

        def work(chunk):
            for p in chunk:
                a, b = off[p], off[p + 1]
                af[a:b] = key_fwd[a:b].astype(np.int64).argsort(kind="quicksort")      # the reference sorts int64
                ar[a:b] = key_rev[a:b].astype(np.int64).argsort(kind="quicksort")

        if self._pool is None or len(pairs) < 2:
            work(pairs)
        else:
            big = [[p] for p in pairs[sizes[pairs] >= 4096]]
            small = pairs[sizes[pairs] < 4096]
            step = max(1, len(small) // (4 * self.sort_threads) + 1)
            chunks = big + [small[i:i + step] for i in range(0, len(small), step)]
            list(self._pool.map(work, chunks))
        return af, ar

    def raw_from_order(self, order):
        """Convert a descending-score permutation (pair-local indices per segment, what
        `sort_values(ascending=False)` yields) into the raw argsort-of-reversed-segment form `extract` takes."""
        order = np.asarray(order)
        out = np.empty(self.n, np.int32)
        off = self.pair_off
        for p in range(self.n_pairs):
            a, b = off[p], off[p + 1]
            out[a:b] = (b - a - 1) - order[a:b][::-1]
        return out

    def extract(self, argsort_fwd, argsort_rev, min_members: int = 1, copy: bool = True) -> Blocks:
        """copy=False returns views of the context's page-locked staging: valid until the next `extract` on this
        engine (what the file writer and the benchmark use); copy=True detaches the arrays."""
        out = icl_blocks()
        af = np.ascontiguousarray(argsort_fwd, np.int32)
        ar = np.ascontiguousarray(argsort_rev, np.int32)
        self._check(self.lib.icl_extract_fields(self.ctx, af.ctypes.data, ar.ctypes.data, int(min_members), FIELDS[self.fields],
                                                C.byref(out)))
        return self._blocks_from(out, copy)

    def _blocks_from(self, out, copy):
        nb, nm = out.n_blocks, out.n_members

        def arr(ptr, n, dt):          # the staging is reused by the next call
            if not ptr:
                return None               # not part of the field set that was asked for (`fields = "writer"`)
            if n == 0:
                return np.zeros(0, dt)
            # (np.frombuffer over the address: half the cost of np.ctypeslib.as_array, which is 3 us per array and step)
            a = np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(C.cast(ptr, C.c_void_p).value), dt, n)
            return a.copy() if copy else a
        res = Blocks(arr(out.pair, nb, np.int32), arr(out.dir, nb, np.int32), arr(out.score, nb, np.int64),
                     arr(out.m, nb, np.int32), arr(out.L, nb, np.int32), arr(out.N, nb, np.int32),
                     arr(out.N1, nb, np.int64), arr(out.L1, nb, np.int32), arr(out.L2, nb, np.int32),
                     arr(out.p_raw, nb, np.float64), arr(out.member_off, nb + 1, np.int64),
                     arr(out.loc1, nm, np.int32), arr(out.loc2, nm, np.int32), arr(out.member_idx, nm, np.int32))
        self.lib.icl_free_blocks(C.byref(out))
        return res

    def debug_dp(self, direction):
        n = self.n
        score, pred, depth = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        used = np.zeros(n, np.uint8)
        self._check(self.lib.icl_debug_dp(self.ctx, int(direction), score.ctypes.data, pred.ctypes.data,
                                          depth.ctypes.data, used.ctypes.data))
        return score, pred, depth, used

    def timings(self):
        ms = np.zeros(len(_lib.ICL_T_NAMES), np.float64)
        self.lib.icl_timings(self.ctx, ms.ctypes.data, len(ms))
        return dict(zip(_lib.ICL_T_NAMES, ms.tolist()))

    def counters(self):
        c = np.zeros(8, np.int64)
        self.lib.icl_counters(self.ctx, c.ctypes.data, 8)
        return dict(edges=int(c[0]), launches_dp=int(c[1]), launches_extract=int(c[2]), rows=int(c[3]),
                    max_span=int(c[4]), chain_kernel={0: "generic", 1: "ring", 3: "single"}[int(c[5])],
                    launches_rank=int(c[6]), cub_calls=int(c[7]))

    # ---- whole hot path ------------------------------------------------------------------------
    def dp_rank_streamed(self, params: ChainParams):
        """`dp` + `rank` overlapped: the DP kernel hands every finished (pair, direction) item to the host through
        page-locked memory (icl_dp_begin) and the sort pool argsorts the segments while heavier pairs are still being
        chained.  Same arrays and same numpy routine as `dp` followed by `rank`."""
        prm = icl_params(int(params.mg1), int(params.mg2), int(params.grading[0]), int(params.over_gap),
                         float(params.coverage_ratio))
        n, P = self.n, self.n_pairs
        keys = self._buf("keys", 2 * n, np.int32)
        both = self._buf("argsort", 2 * n, np.int32)
        flags = self._buf("done", 2 * P, np.int32)
        t0 = time.perf_counter()
        self._check(self.lib.icl_dp_begin(self.ctx, C.byref(prm), keys.ctypes.data, keys.ctypes.data + 4 * n,
                                          flags.ctypes.data))
        self.pair_off = np.zeros(P + 1, np.int64)
        self._check(self.lib.icl_layout(self.ctx, self.pair_off.ctypes.data, None))
        off = self.pair_off
        sizes = np.diff(off)
        off2 = np.concatenate([off, off[1:] + n])
        order = np.argsort(-np.concatenate([sizes, sizes]), kind="stable").astype(np.int32)
        t1 = time.perf_counter()
        try:
            _npsort.seg_argsort(keys, off2, order, both, self.sort_threads, self.flush_keys, flags, 120.0)
        finally:
            self._check(self.lib.icl_dp_end(self.ctx))
        t2 = time.perf_counter()
        self.host_ms = dict(dp=(t1 - t0) * 1e3, rank=(t2 - t1) * 1e3)
        return both[:n], both[n:]

    def numpy_sort_kind(self):
        """Which of the device's exact reproductions of numpy.argsort(kind='quicksort') on int64 matches numpy in THIS
        process: "numpy_avx512" (x86-simd-sort's AVX-512 kernels, hosts with AVX-512), "numpy_avx2" (its AVX2 kernels: x86
        hosts without AVX-512), "numpy_scalar" (numpy's scalar introsort: no SIMD sort dispatched — other CPUs, or
        NPY_DISABLE_CPU_FEATURES) or None (something else, e.g. a future numpy: the host ranks).  Checked once per process and device by sorting tie-heavy score arrays both ways; only on a
        match may the device rank replace the host rank without changing a single tie."""
        if self.device not in _NP_KIND:
            kind = None
            rng = np.random.default_rng(20260)
            sizes = [1, 2, 16, 17, 18, 33, 100, 257, 300, 1000, 4099, 9001]
            n = sum(sizes)
            pair = np.repeat(np.arange(len(sizes), dtype=np.int32), sizes)
            loc1 = np.concatenate([np.arange(1, m + 1) for m in sizes]).astype(np.int32)
            loc2 = rng.integers(1, 40, n).astype(np.int32)            # dense columns: long chains, many tied scores
            grad = rng.choice(np.array([50, 30, 25], np.int32), n)
            probe = IclEngine(self.device, sort_threads=1)
            try:
                probe.upload(pair, loc1, loc2, grad, len(sizes))
                kf, kr = probe.dp(ChainParams(mg1=25, mg2=25, grading=(50, 30, 25)))
                keys = np.concatenate([kf, kr])
                off = np.concatenate([probe.pair_off, probe.pair_off[1:] + n])
                exp = [keys[a:b].astype(np.int64).argsort(kind="quicksort") for a, b in zip(off[:-1], off[1:])]
                got = np.zeros(2 * n, np.int32)
                for name, mode in (("numpy_avx512", 2), ("numpy_avx2", 3), ("numpy_scalar", 1)):
                    try:
                        probe._check(probe.lib.icl_rank_device(probe.ctx, mode))
                        probe._check(probe.lib.icl_debug_order(probe.ctx, got.ctypes.data))
                    except IclError:
                        continue
                    if all((got[a:b] == e).all() for a, b, e in zip(off[:-1], off[1:], exp)) and probe._probe_size_classes(mode):
                        kind = name
                        break
            except IclError:
                kind = None                                    # e.g. a backend without the kernels: the host ranks
            finally:
                probe.close()
            _NP_KIND[self.device] = kind
        return _NP_KIND[self.device]

    # one tie-heavy segment per launch class of the rank kernels (include/icl_b200.h, icl_rank_device: <= 16 000 keys,
    # <= 34 000, <= 65 535, longer) plus the network / small-partition sizes, straight on key arrays (no DP needed)
    PROBE_CLASS_SIZES = (70001, 36007, 17003, 9001, 4099, 257, 256, 33)

    @staticmethod
    def _score_like_keys(rng, m, flavour):
        """Key arrays shaped like DP score segments: a mass of never-updated points holding the three grading values,
        chains of growing scores on top; `flavour` 1 = only the three values (every pass of the SIMD sort stalls and the
        std::sort fallback does the work), 2 = many distinct values."""
        k = rng.choice(np.array([25, 30, 50], np.int32), m, p=[0.5, 0.3, 0.2])
        if flavour == 0:
            chain = rng.random(m) < 0.4
            k[chain] += rng.integers(1, 40, int(chain.sum())).astype(np.int32) * rng.integers(1, 60, int(chain.sum())).astype(np.int32)
        elif flavour == 2:
            k = rng.integers(25, 1 << 20, m).astype(np.int32)
        return k

    def _probe_size_classes(self, mode):
        """True when the device argsort of flavour `mode` equals the running numpy's on every probe segment."""
        rng = np.random.default_rng(4242)
        segs = [self._score_like_keys(rng, m, f) for f in (0, 1, 2) for m in self.PROBE_CLASS_SIZES]
        off = np.r_[0, np.cumsum([len(x) for x in segs])].astype(np.int64)
        keys = np.ascontiguousarray(np.concatenate(segs), np.int32)
        got = np.zeros(len(keys), np.int32)
        try:
            self._check(self.lib.icl_debug_rank_keys(self.ctx, int(mode), len(segs), off.ctypes.data, keys.ctypes.data, got.ctypes.data))
        except IclError:
            return False
        return all((got[a:b] == x.astype(np.int64).argsort(kind="quicksort")).all() for a, b, x in zip(off[:-1], off[1:], segs))

    def _verify_device_rank(self):
        """ICL_B200_RANK_VERIFY: a sample of the segments the device just ranked, re-sorted with numpy in this process."""
        n, P = self.n, self.n_pairs
        if n == 0:
            return True
        keys, order = self._buf("vkeys", 2 * n, np.int32), self._buf("vorder", 2 * n, np.int32)
        self._check(self.lib.icl_debug_keys(self.ctx, keys.ctypes.data))
        self._check(self.lib.icl_debug_order(self.ctx, order.ctypes.data))
        off = np.concatenate([self.pair_off, self.pair_off[1:] + n])
        sizes = np.diff(off)
        cand = np.flatnonzero(sizes > 1)
        if len(cand) == 0:
            return True
        longest = cand[np.argsort(-sizes[cand], kind="stable")[:4]]
        rest = self._verify_rng.choice(cand, min(len(cand), max(0, self.rank_verify - len(longest))), replace=False)
        for sgm in np.unique(np.r_[longest, rest]):
            a, b = off[sgm], off[sgm + 1]
            self.rank_verified_segments += 1
            if not (order[a:b] == keys[a:b].astype(np.int64).argsort(kind="quicksort")).all():
                return False
        return True

    def numpy_sort_is_scalar(self):
        return self.numpy_sort_kind() == "numpy_scalar"

    def run_resident_device_rank(self, params: ChainParams, min_members: int = 1, copy: bool = True,
                                 mode: str = "stable") -> Blocks:
        """Like `run_resident`, but the extraction order comes from the device (icl_rank_device) instead of numpy on the
        host.  mode "stable": NON-REFERENCE tie order (same blocks wherever scores are distinct); mode "numpy_scalar":
        numpy's scalar introsort reproduced exactly — the reference's tie order wherever numpy has no SIMD sort."""
        prm = icl_params(int(params.mg1), int(params.mg2), int(params.grading[0]), int(params.over_gap),
                         float(params.coverage_ratio))
        t0 = time.perf_counter()
        self._check(self.lib.icl_dp(self.ctx, C.byref(prm), None, None))
        self.pair_off = np.zeros(self.n_pairs + 1, np.int64)
        self._check(self.lib.icl_layout(self.ctx, self.pair_off.ctypes.data, None))
        def by_numpy():
            kf, kr = self.dp(params)
            af, ar = self.rank(kf, kr)
            return self.extract(af, ar, min_members, copy)
        out = icl_blocks()
        try:
            self._check(self.lib.icl_rank_device(self.ctx, {"stable": 0, "numpy_scalar": 1, "numpy_avx512": 2, "numpy_avx2": 3}[mode]))
            if self.rank_verify and mode != "stable" and not self._verify_device_rank():
                # the device's permutation is not this numpy's: numpy ranks this comparison and every later one of the process
                self.rank_verify_failures += 1
                _NP_KIND[self.device] = None
                return by_numpy()
            # (the rank kernels report a segment they do not cover — std::sort's heap-sort branch — without stopping the
            # stream: the flag is read at icl_extract's first synchronisation, which then fails with the same code)
            self._check(self.lib.icl_extract_fields(self.ctx, None, None, int(min_members), FIELDS[self.fields], C.byref(out)))
        except IclError as e:
            if e.code != -3 or mode == "stable":
                raise
            self.host_rank_fallbacks += 1                  # numpy itself ranks this comparison
            return by_numpy()
        res = self._blocks_from(out, copy)
        self.host_ms = dict(total=(time.perf_counter() - t0) * 1e3, rank=0.0)
        return res

    def run_resident(self, params: ChainParams, min_members: int = 1, copy: bool = True) -> Blocks:
        """Hot path on points already uploaded: sort + DP -> rank -> extraction + p values.  The rank is numpy's own
        argsort on the host, except when numpy in this process runs its scalar introsort (no SIMD sort dispatched): the
        device reproduces that algorithm exactly, so it ranks there and the host is out of the step (`rank_mode`)."""
        if self.rank_mode == "auto":
            kind = self.numpy_sort_kind()
            if kind:
                return self.run_resident_device_rank(params, min_members, copy, mode=kind)
        elif self.rank_mode in ("numpy_scalar", "numpy_avx512", "numpy_avx2"):
            return self.run_resident_device_rank(params, min_members, copy, mode=self.rank_mode)
        if self.stream_rank and _npsort_ok():
            af, ar = self.dp_rank_streamed(params)
            t2 = time.perf_counter()
            blocks = self.extract(af, ar, min_members, copy)
            self.host_ms["extract"] = (time.perf_counter() - t2) * 1e3
            return blocks
        t0 = time.perf_counter()
        kf, kr = self.dp(params)
        t1 = time.perf_counter()
        af, ar = self.rank(kf, kr)
        t2 = time.perf_counter()
        blocks = self.extract(af, ar, min_members, copy)
        t3 = time.perf_counter()
        self.host_ms = dict(dp=(t1 - t0) * 1e3, rank=(t2 - t1) * 1e3, extract=(t3 - t2) * 1e3)
        return blocks

    def chain(self, pair_id, loc1, loc2, grading, n_pairs, params: ChainParams, min_members: int = 1,
              copy: bool = True) -> Blocks:
        self.upload(pair_id, loc1, loc2, grading, n_pairs)
        return self.run_resident(params, min_members, copy)
