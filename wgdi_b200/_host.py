"""ctypes binding of libicl_host.so (include/icl_host.h): native m8 ingest, grading and file writer.

Host-only C++ (no CUDA).  When the library is not built the driver uses the pandas readers of
`wgdi_b200.base`, which mirror the reference's; both produce identical points and files.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libicl_host.so")
_LIB = None


class icl_hits(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_rows_file", C.c_int64), ("bits_is_int", C.c_int32),
                ("row", C.POINTER(C.c_int64)), ("g1", C.POINTER(C.c_int32)), ("g2", C.POINTER(C.c_int32)),
                ("bits", C.POINTER(C.c_double))]


class icl_collinearity(C.Structure):
    _fields_ = [("n_blocks", C.c_int64), ("n_members", C.c_int64),
                ("id_text", C.c_void_p), ("id_off", C.POINTER(C.c_int64)), ("pv_text", C.c_void_p), ("pv_off", C.POINTER(C.c_int64)),
                ("off", C.POINTER(C.c_int64)), ("g1", C.POINTER(C.c_int32)), ("g2", C.POINTER(C.c_int32)),
                ("loc1", C.POINTER(C.c_double)), ("loc2", C.POINTER(C.c_double)),
                ("block1_text", C.c_void_p), ("block1_off", C.POINTER(C.c_int64)),
                ("block2_text", C.c_void_p), ("block2_off", C.POINTER(C.c_int64))]


class icl_ingest_info(C.Structure):
    _fields_ = [("n_hits", C.c_int64), ("n_rows_file", C.c_int64), ("bits_is_int", C.c_int32),
                ("n_amb_queries", C.c_int64), ("n_amb_hits", C.c_int64)]


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        lib.icl_host_read_m8.argtypes = [C.c_char_p, i32, dbl, dbl, vp, vp, i64, vp, vp, i64, i32, C.POINTER(icl_hits), vp, i32]
        lib.icl_host_free_hits.argtypes = [C.POINTER(icl_hits)]
        lib.icl_host_free_hits.restype = None
        lib.icl_host_grade.argtypes = [i64, vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, i32]
        lib.icl_host_rank.argtypes = [i64, vp, vp, i32, i32, i32, vp, vp, i32, vp, i32]
        lib.icl_host_write_blocks.argtypes = [C.c_char_p, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp,
                                              vp, i64, vp, vp, vp, i32, vp, i32]
        lib.icl_host_write_collinearity.argtypes = [C.c_char_p, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                                    vp, i64, vp, vp, vp, vp, i64, vp, vp, vp, i32, vp, i32]
        lib.icl_host_read_collinearity.argtypes = [C.c_char_p, vp, vp, i64, vp, vp, i64, C.POINTER(icl_collinearity), vp, i32]
        lib.icl_host_free_collinearity.argtypes = [C.POINTER(icl_collinearity)]
        lib.icl_host_free_collinearity.restype = None
        lib.icl_host_ingest_open.argtypes = [C.c_char_p, i32, dbl, dbl, vp, vp, i64, vp, vp, i64, i32, i32, i32, i32,
                                             C.POINTER(vp), C.POINTER(icl_ingest_info), vp, i32]
        lib.icl_host_ingest_ambiguous.argtypes = [vp, vp, vp]
        lib.icl_host_ingest_set_classes.argtypes = [vp, vp]
        lib.icl_host_ingest_count.argtypes = [vp, vp, vp, i32, i32, vp, C.POINTER(i64)]
        lib.icl_host_ingest_points.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
        lib.icl_host_ingest_close.argtypes = [vp]
        lib.icl_host_ingest_close.restype = None
        lib.icl_host_place_blocks.argtypes = [i64, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
        _LIB = lib
    return _LIB


def place_blocks(pair, global_pair, block_base, member_base, member_off, block_fields, member_fields, out_pair, out_member_off,
                 member_idx=None, rows=None, out_member_idx=None):
    """icl_host_place_blocks: block_fields / member_fields = lists of (source array, destination array) of equal dtype."""
    lib = load()

    def table(fields):
        n = len(fields)
        src = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a, _ in fields])
        dst = (C.c_void_p * max(n, 1))(*[d.ctypes.data for _, d in fields])
        wid = (C.c_int32 * max(n, 1))(*[a.dtype.itemsize for a, _ in fields])
        for a, d in fields:
            if a.dtype != d.dtype or not a.flags.c_contiguous:
                raise ValueError("place_blocks: source and destination arrays must be contiguous and of one dtype")
        return n, src, dst, wid
    nbf, bs, bd, bw = table(block_fields)
    nmf, ms, md, mw = table(member_fields)
    pair = np.ascontiguousarray(pair, np.int32)
    global_pair = np.ascontiguousarray(global_pair, np.int32)
    block_base = np.ascontiguousarray(block_base, np.int64)
    member_base = np.ascontiguousarray(member_base, np.int64)
    member_off = np.ascontiguousarray(member_off, np.int64)
    with_idx = member_idx is not None and rows is not None and out_member_idx is not None
    if with_idx:
        member_idx = np.ascontiguousarray(member_idx, np.int32)
        rows = np.ascontiguousarray(rows, np.int64)
    rc = lib.icl_host_place_blocks(len(pair), pair.ctypes.data, global_pair.ctypes.data, block_base.ctypes.data,
                                   member_base.ctypes.data, member_off.ctypes.data, nbf, bs, bd, bw, nmf, ms, md, mw,
                                   out_pair.ctypes.data, out_member_off.ctypes.data,
                                   member_idx.ctypes.data if with_idx else None, rows.ctypes.data if with_idx else None,
                                   out_member_idx.ctypes.data if with_idx else None)
    if rc:
        raise RuntimeError(f"icl_host_place_blocks failed ({rc})")


_COMPRESSED_EXT = (".gz", ".bz2", ".zip", ".xz", ".zst", ".tar", ".tar.gz", ".tar.xz", ".tar.bz2")
_COMPRESSED_MAGIC = (b"\x1f\x8b", b"BZh", b"\xfd7zXZ\x00", b"PK\x03\x04", b"\x28\xb5\x2f\xfd")


def is_compressed(path):
    """True when pandas.read_csv would (have to) decompress `path`: the reference reads the BLAST table with pandas
    (wgdi/base.py:174), which infers gzip / bz2 / zip / xz / zstd from the extension.  The native parser maps raw bytes, so
    such files go through the pandas reader (also when only the magic bytes say so: then pandas fails as the reference does)."""
    if str(path).lower().endswith(_COMPRESSED_EXT):
        return True
    try:
        with open(path, "rb") as f:
            head = f.read(6)
    except OSError:
        return False
    return head.startswith(_COMPRESSED_MAGIC)


class HostError(RuntimeError):
    pass


def _call(fn, *args):
    err = C.create_string_buffer(512)
    rc = fn(*args, err, 512)
    if rc != 0:
        raise HostError(f"icl_host error {rc}: {err.value.decode(errors='replace')}")


def pack_strings(values):
    """(bytes, offsets[n+1]) of a sequence of str."""
    vals = [str(v) for v in (values.tolist() if hasattr(values, "tolist") else values)]
    off = np.zeros(len(vals) + 1, np.int64)
    joined = "".join(vals).encode()
    np.cumsum(np.fromiter(map(len, vals), np.int64, len(vals)), out=off[1:])
    if int(off[-1]) == len(joined):                       # ASCII ids: characters are bytes
        return joined, off
    enc = [v.encode() for v in vals]
    np.cumsum([len(b) for b in enc], out=off[1:])
    return b"".join(enc), off


def threads_default():
    """Host threads of the native calls: the cores this process may run on (cgroup / affinity aware), at most 32."""
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    return max(1, min(32, n))


def read_m8(path, reverse, min_score, max_evalue, ids1, ids2, threads=None):
    """Rows of the m8 table that survive wgdi/base.py:176-180.  ids1/ids2: (bytes, offsets) gene tables.
    Returns dict(row, g1, g2, bits) with bits typed like pandas would infer column 11."""
    lib = load()
    out = icl_hits()
    b1, o1 = ids1
    b2, o2 = ids2
    same = ids1 is ids2
    p1 = C.cast(C.c_char_p(b1), C.c_void_p)
    p2 = p1 if same else C.cast(C.c_char_p(b2), C.c_void_p)
    _call(lib.icl_host_read_m8, str(path).encode(), 1 if reverse else 0, float(min_score), float(max_evalue),
          p1, o1.ctypes.data, len(o1) - 1, p2, (o1 if same else o2).ctypes.data, len(o2) - 1,
          int(threads or threads_default()), C.byref(out))
    try:
        n = out.n

        def arr(ptr, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)
        bits = arr(out.bits, np.float64)
        if out.bits_is_int:
            bits = bits.astype(np.int64)
        return dict(row=arr(out.row, np.int64), g1=arr(out.g1, np.int32), g2=arr(out.g2, np.int32), bits=bits,
                    n_rows_file=int(out.n_rows_file))
    finally:
        lib.icl_host_free_hits(C.byref(out))


class Ingest:
    """The fused ingest of `comparison = genomes` (icl_host_ingest_*): m8 table -> graded points with nothing per hit
    handled in Python.  Use as a context manager.

        with Ingest(path, reverse, score, evalue, ids1, ids2, multiple, repeat_number) as ing:
            off, bits = ing.ambiguous()          # queries whose classes depend on numpy's tie order (hazard H2)
            ing.set_classes(cls)                 # the caller's numpy ranking of exactly those
            present, n = ing.count(chr1, chr2, nc1, nc2)
            pts = ing.points(order1, order2, nc2, pair_rank, grading)
    """

    def __init__(self, path, reverse, min_score, max_evalue, ids1, ids2, multiple, repeat_number, blue_extra=4, threads=None):
        lib = load()
        self._lib = lib
        self._h = C.c_void_p()
        self.info = icl_ingest_info()
        b1, o1 = ids1
        b2, o2 = ids2
        same = ids1 is ids2
        p1 = C.cast(C.c_char_p(b1), C.c_void_p)
        p2 = p1 if same else C.cast(C.c_char_p(b2), C.c_void_p)
        _call(lib.icl_host_ingest_open, str(path).encode(), 1 if reverse else 0, float(min_score), float(max_evalue),
              p1, o1.ctypes.data, len(o1) - 1, p2, (o1 if same else o2).ctypes.data, len(o2) - 1,
              int(multiple), int(repeat_number), int(blue_extra), int(threads or threads_default()),
              C.byref(self._h), C.byref(self.info))
        self.n_hits = int(self.info.n_hits)
        self._n_points = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self._h:
            self._lib.icl_host_ingest_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ambiguous(self):
        """(offsets[n_queries + 1], bit scores typed like pandas would infer column 11) of the queries left ungraded."""
        off = np.zeros(int(self.info.n_amb_queries) + 1, np.int64)
        bits = np.zeros(int(self.info.n_amb_hits), np.float64)
        if self._lib.icl_host_ingest_ambiguous(self._h, off.ctypes.data, bits.ctypes.data):
            raise HostError("icl_host_ingest_ambiguous failed")
        return off, (bits.astype(np.int64) if self.info.bits_is_int else bits)

    def set_classes(self, cls):
        cls = np.ascontiguousarray(cls, np.int8)
        if len(cls) != int(self.info.n_amb_hits) or self._lib.icl_host_ingest_set_classes(self._h, cls.ctypes.data):
            raise HostError("icl_host_ingest_set_classes: one class in 0..3 per ambiguous hit expected")

    def count(self, chr1, chr2, nc1, nc2):
        self._chr = (np.ascontiguousarray(chr1, np.int32), np.ascontiguousarray(chr2, np.int32))
        present = np.zeros(max(int(nc1), 1) * max(int(nc2), 1), np.uint8)
        n = C.c_int64(0)
        rc = self._lib.icl_host_ingest_count(self._h, self._chr[0].ctypes.data, self._chr[1].ctypes.data, max(int(nc1), 1),
                                             max(int(nc2), 1), present.ctypes.data, C.byref(n))
        if rc:
            raise HostError(f"icl_host_ingest_count failed ({rc})")
        self._n_points = int(n.value)
        return present, self._n_points

    def points(self, order1, order2, nc2, pair_rank, grading):
        """dict(pair_id, loc1, loc2, grading, grade_class), int32, file order — after `count`."""
        n = self._n_points
        o1 = np.ascontiguousarray(order1, np.int64)
        o2 = np.ascontiguousarray(order2, np.int64)
        rank = np.ascontiguousarray(pair_rank, np.int32)
        g3 = np.ascontiguousarray(list(grading)[:3], np.int32)
        out = {k: np.empty(n, np.int32) for k in ("pair_id", "loc1", "loc2", "grading", "grade_class")}
        rc = self._lib.icl_host_ingest_points(self._h, self._chr[0].ctypes.data, o1.ctypes.data, self._chr[1].ctypes.data,
                                              o2.ctypes.data, max(int(nc2), 1), rank.ctypes.data, g3.ctypes.data,
                                              *[out[k].ctypes.data for k in ("pair_id", "loc1", "loc2", "grading", "grade_class")])
        if rc:
            raise HostError(f"icl_host_ingest_points failed ({rc})")
        return out


def grade(query, bits, n_query, multiple, repeat_number, grading, threads=None):
    """(grading per hit, ambiguous flag per query) — see icl_host_grade."""
    lib = load()
    q = np.ascontiguousarray(query, np.int32)
    b = np.ascontiguousarray(bits, np.float64)
    g3 = np.ascontiguousarray(grading[:3], np.int32)
    out = np.zeros(len(q), np.int32)
    amb = np.zeros(max(int(n_query), 1), np.uint8)
    _call(lib.icl_host_grade, len(q), q.ctypes.data, b.ctypes.data, int(n_query), int(multiple), int(repeat_number),
          g3.ctypes.data, out.ctypes.data, amb.ctypes.data, int(threads or threads_default()))
    return out, amb[:n_query].astype(bool)


def rank(query, bits, n_query, repeat_number, class_bound, threads=None):
    """(rank per hit, ambiguous flag per query) — see icl_host_rank."""
    lib = load()
    q = np.ascontiguousarray(query, np.int32)
    b = np.ascontiguousarray(bits, np.float64)
    out = np.full(len(q), -1, np.int64)
    amb = np.zeros(max(int(n_query), 1), np.uint8)
    _call(lib.icl_host_rank, len(q), q.ctypes.data, b.ctypes.data, int(n_query), int(repeat_number), int(class_bound),
          out.ctypes.data, amb.ctypes.data, int(threads or threads_default()))
    return out, amb[:n_query].astype(bool)


def write_blocks(path, headers, mem_begin, mem_end, loc1, loc2, chr1_code, chr2_code, genes1, genes2, threads=None):
    lib = load()
    hb, ho = pack_strings(headers)
    a = [np.ascontiguousarray(x, dt) for x, dt in ((mem_begin, np.int64), (mem_end, np.int64), (loc1, np.int32),
                                                  (loc2, np.int32), (chr1_code, np.int32), (chr2_code, np.int32))]
    _call(lib.icl_host_write_blocks, str(path).encode(), len(headers), C.cast(C.c_char_p(hb), C.c_void_p), ho.ctypes.data,
          *[x.ctypes.data for x in a],
          genes1["row_of"].ctypes.data, int(genes1["maxo"]), C.cast(C.c_char_p(genes1["name_bytes"]), C.c_void_p),
          genes1["name_off"].ctypes.data, genes1["strand_code"].ctypes.data,
          genes2["row_of"].ctypes.data, int(genes2["maxo"]), C.cast(C.c_char_p(genes2["name_bytes"]), C.c_void_p),
          genes2["name_off"].ctypes.data, genes2["strand_code"].ctypes.data,
          int(threads or threads_default()))


def write_collinearity(path, score, p_text, p_index, pair_text, pair, mem_begin, mem_end, loc1, loc2, chr1_code, chr2_code,
                       genes1, genes2, threads=None):
    """icl_host_write_collinearity: headers and body of the collinearity file from arrays (one entry per emitted block)."""
    lib = load()
    pb, po = pack_strings(p_text)
    nb, no = pack_strings(pair_text)
    a = [np.ascontiguousarray(x, dt) for x, dt in ((score, np.int64), (p_index, np.int32), (pair, np.int32),
                                                  (mem_begin, np.int64), (mem_end, np.int64), (loc1, np.int32),
                                                  (loc2, np.int32), (chr1_code, np.int32), (chr2_code, np.int32))]
    sc, pi, pr, mb, me, l1, l2, c1, c2 = a
    _call(lib.icl_host_write_collinearity, str(path).encode(), len(sc), sc.ctypes.data,
          C.cast(C.c_char_p(pb), C.c_void_p), po.ctypes.data, pi.ctypes.data,
          C.cast(C.c_char_p(nb), C.c_void_p), no.ctypes.data, pr.ctypes.data,
          mb.ctypes.data, me.ctypes.data, l1.ctypes.data, l2.ctypes.data, c1.ctypes.data, c2.ctypes.data,
          genes1["row_of"].ctypes.data, int(genes1["maxo"]), C.cast(C.c_char_p(genes1["name_bytes"]), C.c_void_p),
          genes1["name_off"].ctypes.data, genes1["strand_code"].ctypes.data,
          genes2["row_of"].ctypes.data, int(genes2["maxo"]), C.cast(C.c_char_p(genes2["name_bytes"]), C.c_void_p),
          genes2["name_off"].ctypes.data, genes2["strand_code"].ctypes.data,
          int(threads or threads_default()))


def read_collinearity(path, ids1, ids2):
    """icl_host_read_collinearity: dict(ids, pvalues, off, g1, g2, loc1, loc2, block1, block2) of a WGDI collinearity file, or
    None when the file is not the plain case the native reader handles (the caller's Python reader takes over)."""
    lib = load()
    out = icl_collinearity()
    b1, o1 = ids1
    b2, o2 = ids2
    same = ids1 is ids2
    p1 = C.cast(C.c_char_p(b1), C.c_void_p)
    p2 = p1 if same else C.cast(C.c_char_p(b2), C.c_void_p)
    err = C.create_string_buffer(512)
    rc = lib.icl_host_read_collinearity(str(path).encode(), p1, o1.ctypes.data, len(o1) - 1, p2,
                                        (o1 if same else o2).ctypes.data, len(o2) - 1, C.byref(out), err, 512)
    if rc == 1:
        return None
    if rc != 0:
        raise HostError(f"icl_host error {rc}: {err.value.decode(errors='replace')}")
    try:
        nb, nm = int(out.n_blocks), int(out.n_members)

        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)

        def texts(ptr, offp):
            off = arr(offp, nb + 1, np.int64)
            raw = C.string_at(ptr, int(off[-1])).decode("ascii") if nb else ""
            o = off.tolist()
            return [raw[o[i]:o[i + 1]] for i in range(nb)]
        return dict(ids=texts(out.id_text, out.id_off), pvalues=texts(out.pv_text, out.pv_off),
                    off=arr(out.off, nb + 1, np.int64), g1=arr(out.g1, nm, np.int64), g2=arr(out.g2, nm, np.int64),
                    loc1=arr(out.loc1, nm, np.float64), loc2=arr(out.loc2, nm, np.float64),
                    block1=texts(out.block1_text, out.block1_off), block2=texts(out.block2_text, out.block2_off))
    finally:
        lib.icl_host_free_collinearity(C.byref(out))
