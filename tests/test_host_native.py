"""Native host helpers (libicl_host.so: m8 ingest, grading, writer) against the pandas readers that mirror the
reference and against the oracle.  CPU only."""
import contextlib
import ctypes
import io
import os
import re

import numpy as np
import pytest

from conftest import ROOT, stage_e2e
from oracle import icl_oracle


@pytest.fixture(scope="module")
def host():
    from wgdi_b200 import _host, build
    build.build_host()
    _host._LIB = None
    return _host


def test_host_library_exports_declared_symbols(host):
    lib = ctypes.CDLL(host.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "icl_host.h")).read()
    declared = set(re.findall(r"\b(icl_host_[a-z_0-9]+)\s*\(", header))
    assert declared == {"icl_host_read_m8", "icl_host_free_hits", "icl_host_grade", "icl_host_rank", "icl_host_write_blocks", "icl_host_place_blocks",
                        "icl_host_write_collinearity", "icl_host_read_collinearity", "icl_host_free_collinearity",
                        "icl_host_ingest_open", "icl_host_ingest_ambiguous", "icl_host_ingest_set_classes",
                        "icl_host_ingest_count", "icl_host_ingest_points", "icl_host_ingest_close"}
    for name in declared:
        assert hasattr(lib, name)


def _write(path, rows):
    with open(path, "w", newline="") as f:
        f.write(rows)


def test_read_m8_filters_like_newblast(host, tmp_path):
    import pandas as pd
    from wgdi_b200 import base
    genes = [f"g{i}" for i in range(8)]
    gff = pd.DataFrame({"chr": ["1"] * 8}, index=pd.Index(genes))
    rows = [
        ("g0", "g1", "1e-30", "250"), ("g0", "g1", "1e-20", "200"),      # duplicate pair: first kept
        ("g0", "g0", "0", "900"),                                       # self hit
        ("g1", "g2", "1.00e-05", "300"),                                # evalue == threshold: strict <
        ("g1", "g3", "9.99e-06", "100"),                                # score == threshold: >=
        ("g1", "g4", "1e-50", "99"), ("zz", "g1", "1e-50", "500"), ("g1", "zz", "1e-50", "500"),
        ("g2", "g5", "3e-12", "120"), ("g7", "g6", "2e-9", "101"),
    ]
    def table(score_fmt=lambda s: s, sep="\n", blank=False, extra=False):
        out = []
        for a, b, ev, sc in rows:
            cols = [a, b, "90.1", "100", "1", "0", "1", "100", "1", "100", ev, score_fmt(sc)] + (["x"] if extra else [])
            out.append("\t".join(cols))
            if blank:
                out.append("")
        return sep.join(out) + sep
    ids = host.pack_strings(genes)
    for kind, text in (("int", table()), ("crlf+blank", table(sep="\r\n", blank=True)),
                       ("float", table(score_fmt=lambda s: s + ".5" if s == "120" else s))):
        p = tmp_path / f"{kind}.m8"
        _write(p, text)
        got = host.read_m8(str(p), False, 100, 1e-5, ids, ids, threads=3)
        exp = base.newblast(str(p), 100, 1e-5, gff, gff, "false")
        assert (got["g1"] == exp["g1"].to_numpy()).all() and (got["g2"] == exp["g2"].to_numpy()).all(), kind
        assert got["bits"].dtype == exp[11].to_numpy().dtype and (got["bits"] == exp[11].to_numpy()).all(), kind
        assert (got["row"] == exp.index.to_numpy()).all(), kind
    # reverse swaps the id columns first
    p = tmp_path / "rev.m8"
    _write(p, table())
    got = host.read_m8(str(p), True, 100, 1e-5, ids, ids)
    exp = base.newblast(str(p), 100, 1e-5, gff, gff, "true")
    assert (got["g1"] == exp["g1"].to_numpy()).all() and (got["g2"] == exp["g2"].to_numpy()).all()
    # malformed rows are an error, a missing file too
    _write(tmp_path / "bad.m8", "g0\tg1\t1\n")
    with pytest.raises(host.HostError):
        host.read_m8(str(tmp_path / "bad.m8"), False, 100, 1e-5, ids, ids)
    with pytest.raises(host.HostError):
        host.read_m8(str(tmp_path / "missing.m8"), False, 100, 1e-5, ids, ids)
    # empty file
    _write(tmp_path / "empty.m8", "")
    assert len(host.read_m8(str(tmp_path / "empty.m8"), False, 100, 1e-5, ids, ids)["g1"]) == 0


def test_read_m8_number_forms_near_the_thresholds(host, tmp_path):
    """The reader decides `evalue < limit` from an extended-precision estimate unless the token is within a hair of
    the limit, and converts bit scores on a fast exact path: both must agree with pandas' parser for every spelling."""
    import pandas as pd
    from wgdi_b200 import base
    # tokens of at most 15 significant digits: beyond that pandas' default parser is itself off by an ulp now and then
    # ("9.999999999999999e-06" reads as 1e-05) and BLAST never prints such numbers
    evs = ["1e-5", "1E-05", "0.00001", "1.00000000000001e-05", "9.99999999999999e-06", "0.999999999999999e-5",
           "99999999999999e-19", "100000000000001e-19", "1e-400", "0", "0.0", "0e0", "-1e-3", "+2e-7",
           "1e-05", "9.9e-6", "1.1e-5", "1e+2", "123456789012345e-27", ".5e-5", "5.e-6", "7e-006", "1.00e-300", "4.2e-161"]
    scs = ["100", "99.9999999999999", "100.000000000001", "1e2", "100.0", "0100", "+250", "99", "1234.5", "59.3",
           "100.5e0", "123456789012345e-12", "1e1", "1000000", "2.5E+2", "99.5"]
    n = len(evs) * len(scs)
    genes = [f"g{i}" for i in range(n + 1)]
    gff = pd.DataFrame({"chr": ["1"] * len(genes)}, index=pd.Index(genes))
    lines, k = [], 0
    for ev in evs:
        for sc in scs:
            k += 1
            lines.append("\t".join(["g0", f"g{k}", "90.1", "100", "1", "0", "1", "100", "1", "100", ev, sc]))
    p = tmp_path / "forms.m8"
    _write(p, "\n".join(lines) + "\n")
    ids = host.pack_strings(genes)
    for limit, score in ((1e-5, 100), (1e-05, 99.9999999999999), (9.99999999999999e-06, 100), (4.2e-7, 59.3)):
        got = host.read_m8(str(p), False, score, limit, ids, ids, threads=2)
        exp = base.newblast(str(p), score, limit, gff, gff, "false")
        assert (got["row"] == exp.index.to_numpy()).all(), (limit, score)
        assert (got["bits"] == exp[11].to_numpy()).all()
        assert 0 < len(got["row"]) < n


def test_native_grading_matches_oracle_with_ties(host):
    from wgdi_b200.ingest import grade_hits_native
    rng = np.random.default_rng(21)
    for trial in range(150):
        n, nq = int(rng.integers(1, 500)), int(rng.integers(1, 40))
        q = rng.integers(0, nq, n).astype(np.int32)
        bits = rng.integers(100, 104 + trial % 7 * 20, n).astype(np.int64 if trial % 3 else np.float64)
        mult, rn = int(rng.integers(0, 4)), int(rng.integers(1, 25))
        got = grade_hits_native(q, bits, nq, mult, rn, [50, 30, 25])
        assert (got == icl_oracle.grade_hits(q, bits, mult, rn, [50, 30, 25])).all()


def _random_m8(rng, path, genes, n_rows, sorted_by_query, tie_span, float_bits=False, crlf=False):
    """A BLAST-like table with duplicate pairs, self hits, unknown ids, sub-threshold rows and many tied bit scores."""
    q = rng.integers(0, len(genes), n_rows)
    if sorted_by_query:
        q = np.sort(q)
    s = rng.integers(0, len(genes) + 3, n_rows)                     # a few subjects are not in the gene table
    bits = rng.integers(95, 100 + tie_span, n_rows)
    ev = rng.choice(["1e-30", "2.5e-7", "0.0", "1e-5", "3e-4", "9.9e-06"], n_rows, p=[0.5, 0.2, 0.1, 0.05, 0.1, 0.05])
    names = list(genes) + ["zz1", "zz2", "zz3"]
    eol = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        for k, (a, b, e, sc) in enumerate(zip(q, s, ev, bits)):
            sc = f"{sc}.5" if float_bits and k % 3 == 0 else str(sc)
            f.write("\t".join([genes[a], names[b], "90.1", "100", "1", "0", "1", "100", "1", "100", e, sc]) + eol)
            if crlf and k % 50 == 0:
                f.write(eol)                                        # blank lines are skipped (pandas: skip_blank_lines)


@pytest.mark.parametrize("sorted_by_query", [True, False])
def test_fused_ingest_equals_the_pandas_readers(host, tmp_path, sorted_by_query):
    """icl_host_ingest_* (parse + filters + duplicates + grading + point table in one handle) against the reference-like
    pandas path and the oracle's grading, on tables full of boundary ties: query-grouped like BLAST writes them, and
    shuffled (the grouping by query must not depend on the file order)."""
    import pandas as pd
    from wgdi_b200 import base
    from wgdi_b200.ingest import classes_of_ambiguous, pair_ids, pair_table
    rng = np.random.default_rng(5 + sorted_by_query)
    n_amb = 0
    for trial in range(12):
        ng = int(rng.integers(3, 60))
        genes = [f"gene{i:03d}_{'x' * int(rng.integers(0, 30))}" for i in range(ng)]     # ids longer than a map slot too
        chrom = rng.integers(0, 4, ng)
        order = rng.integers(1, 50, ng).astype(np.int64)
        gff = pd.DataFrame({"chr": [f"c{c}" for c in chrom]}, index=pd.Index(genes))
        p = tmp_path / f"t{trial}.m8"
        float_bits, crlf, reverse = trial % 3 == 1, trial % 4 == 2, trial % 5 == 3
        _random_m8(rng, p, genes, int(rng.integers(1, 3000)), sorted_by_query, tie_span=int(rng.integers(2, 40)),
                   float_bits=float_bits, crlf=crlf)
        mult, rn = int(rng.integers(1, 4)), int(rng.integers(2, 30))
        grading = [50, 30, 25]
        # reference-like path
        blast = base.newblast(str(p), 100, 1e-5, gff, gff, "true" if reverse else "false")
        g1, g2, bits = blast["g1"].to_numpy(), blast["g2"].to_numpy(), blast[11].to_numpy()
        grade = icl_oracle.grade_hits(g1, bits, mult, rn, [1, 2, 3])
        ok = grade > 0
        chr_names, chr_code = np.unique(np.array([f"c{c}" for c in chrom]), return_inverse=True)
        pid, names = pair_ids(list(chr_names), list(chr_names), chr_code[g1[ok]], chr_code[g2[ok]])
        # fused path
        ids = host.pack_strings(genes)
        with host.Ingest(str(p), reverse, 100, 1e-5, ids, ids, mult, rn, threads=int(rng.integers(1, 6))) as ing:
            assert ing.n_hits == len(g1)
            off, abits = ing.ambiguous()
            assert abits.dtype == bits.dtype
            n_amb += len(off) - 1
            ing.set_classes(classes_of_ambiguous(off, abits, mult, rn))
            present, n = ing.count(chr_code, chr_code, len(chr_names), len(chr_names))
            assert n == int(ok.sum())
            table, names2 = pair_table(list(chr_names), list(chr_names), present)
            pts = ing.points(order, order, len(chr_names), table, grading)
        assert names2 == names
        assert (pts["pair_id"] == pid).all() and (pts["loc1"] == order[g1[ok]]).all() and (pts["loc2"] == order[g2[ok]]).all()
        assert (pts["grade_class"] == grade[ok] - 1).all() and (pts["grading"] == np.asarray(grading)[grade[ok] - 1]).all()
    assert n_amb > 20                                   # the numpy re-ranking was exercised


def test_fused_ingest_errors_and_empty_inputs(host, tmp_path):
    ids = host.pack_strings(["g0", "g1"])
    _write(tmp_path / "bad.m8", "g0\tg1\t1\n")
    with pytest.raises(host.HostError):
        host.Ingest(str(tmp_path / "bad.m8"), False, 100, 1e-5, ids, ids, 1, 5)
    with pytest.raises(host.HostError):
        host.Ingest(str(tmp_path / "missing.m8"), False, 100, 1e-5, ids, ids, 1, 5)
    _write(tmp_path / "empty.m8", "")
    with host.Ingest(str(tmp_path / "empty.m8"), False, 100, 1e-5, ids, ids, 1, 5) as ing:
        assert ing.n_hits == 0
        off, bits = ing.ambiguous()
        assert len(off) == 1 and len(bits) == 0
        ing.set_classes(np.zeros(0, np.int8))
        present, n = ing.count(np.zeros(2, np.int32), np.zeros(2, np.int32), 1, 1)
        assert n == 0 and not present.any()
        pts = ing.points(np.ones(2, np.int64), np.ones(2, np.int64), 1, np.zeros(1, np.int32), [50, 30, 25])
        assert all(len(v) == 0 for v in pts.values())


def test_native_writer_equals_format_blocks(host, tmp_path):
    """icl_host_write_collinearity (headers and body assembled natively) against the Python formatter on synthetic blocks:
    p values of every printed form, the pvalue / over_gap filters, and a chromosome name with a ':' (the reference
    re-cuts the finished header at ':', run_colliearity.py:147-151)."""
    import pandas as pd
    from types import SimpleNamespace
    from wgdi_b200.run_colliearity import _gene_table, format_blocks, write_blocks_native
    rng = np.random.default_rng(11)
    def genome(chrs, per):
        rows = [(c, f"{c.replace(':', '_')}g{i}", i + 1, "+-"[int(rng.integers(0, 2))]) for c in chrs for i in range(per)]
        return pd.DataFrame({"chr": [r[0] for r in rows], "order": [r[2] for r in rows], "strand": [r[3] for r in rows]},
                            index=pd.Index([r[1] for r in rows]))
    g1, g2 = _gene_table(genome(["1", "2", "x:y"], 40)), _gene_table(genome(["A", "B:1"], 40))
    pair_names = sorted((a, b) for a in g1["chr_names"] for b in g2["chr_names"])
    nb = 60
    m = rng.integers(2, 12, nb)
    off = np.zeros(nb + 1, np.int64)
    np.cumsum(m, out=off[1:])
    blocks = SimpleNamespace(pair=np.sort(rng.integers(0, len(pair_names), nb)).astype(np.int32), score=rng.integers(100, 5000, nb),
                             m=m.astype(np.int32), member_off=off,
                             loc1=rng.integers(1, 41, off[-1]).astype(np.int32), loc2=rng.integers(1, 41, off[-1]).astype(np.int32),
                             p_raw=rng.choice([0.0, 1e-9, 0.00004, 0.00005, 0.0001, 0.01234, 0.2, 0.99996, 1.0, 3.2], nb))
    blocks.p_value = lambda: np.round(blocks.p_raw, 4)
    for pvalue, over_gap in ((1.0, 2), (0.05, 5), (0.0, 2)):
        text = format_blocks(blocks, pair_names, g1, g2, pvalue, over_gap)
        path = tmp_path / f"w_{pvalue}_{over_gap}.txt"
        write_blocks_native(str(path), blocks, pair_names, g1, g2, pvalue, over_gap)
        assert open(path).read() == text
    assert "x" in text and ":y" not in text.replace("# Alignment", "")       # the ':' cut happened


@pytest.mark.parametrize("case", ["tiny", "cross", "tiny_rev", "cross_chrom"])
def test_native_and_pandas_paths_write_the_same_file(host, emu_backend, case, tmp_path):
    """conf in -> file out through the native parser / writer and through the pandas readers (fake device backend)."""
    from wgdi_b200 import base
    from wgdi_b200.run_colliearity import mycollinearity
    stage_e2e(case, str(tmp_path))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        texts = {}
        for mode in ("native", "native-steps", "pandas"):
            opts = [(k, v if k != "savefile" else f"{mode}.txt") for k, v in base.load_conf("icl.conf", "collinearity")]
            with contextlib.redirect_stdout(io.StringIO()):
                m = mycollinearity(opts + [("ingest", mode)])
                assert m.native == (mode != "pandas")
                with pytest.raises(SystemExit):
                    m.run()
            texts[mode] = open(f"{mode}.txt").read()
        assert texts["native"] == texts["pandas"] == texts["native-steps"] and len(texts["native"]) > 1000
    finally:
        os.chdir(cwd)


def test_native_place_blocks_equals_numpy(host, monkeypatch):
    """icl_host_place_blocks against the numpy placement of shard.place_blocks: two ranks with interleaved pairs, with and
    without the optional arrays (field set "writer")."""
    from wgdi_b200 import shard
    from wgdi_b200.engine import Blocks
    rng = np.random.default_rng(3)
    P = 40
    owner = rng.integers(0, 2, P)
    parts, rows_of, gps = [], [], []
    for r in range(2):
        gp = np.flatnonzero(owner == r).astype(np.int32)
        nb = 300
        pair = np.sort(rng.integers(0, len(gp), nb)).astype(np.int32)
        m = rng.integers(1, 30, nb).astype(np.int32)
        mo = np.r_[0, np.cumsum(m)].astype(np.int64)
        nm = int(mo[-1])
        rows = rng.permutation(5000)[:2000].astype(np.int64)
        f = lambda dt, n: rng.integers(0, 1000, n).astype(dt)
        parts.append(Blocks(pair, f(np.int32, nb), f(np.int64, nb), m, f(np.int32, nb), f(np.int32, nb), f(np.int64, nb),
                            f(np.int32, nb), f(np.int32, nb), rng.random(nb), mo, f(np.int32, nm), f(np.int32, nm),
                            rng.integers(0, 2000, nm).astype(np.int32)))
        rows_of.append(rows)
        gps.append(gp)
    for lean in (False, True):
        use = [Blocks(*[None if (lean and f in ("dir", "L", "N", "N1", "L1", "L2", "member_idx")) else getattr(b, f)
                        for f in shard.BLOCK_FIELDS]) for b in parts]
        counts = [shard.pair_counts(b, gp, P) for b, gp in zip(use, gps)]
        bb = np.r_[0, np.cumsum(sum(c[0] for c in counts))]
        mb = np.r_[0, np.cumsum(sum(c[1] for c in counts))]
        outs = []
        for native in (True, False):
            monkeypatch.setattr(host, "available", lambda native=native: native)
            out = {f: np.zeros(int(mb[-1]) + 1 if f in shard.MEMBER_FIELDS else int(bb[-1]) + 2, shard.HostGather.DT[f])
                   for f in shard.BLOCK_FIELDS}
            for b, gp, rows in zip(use, gps, rows_of):
                shard.place_blocks(b, gp, rows, bb, mb, out)
            outs.append(out)
        for f in shard.BLOCK_FIELDS:
            assert (outs[0][f] == outs[1][f]).all(), (lean, f)
