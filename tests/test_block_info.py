"""`-bi` (block information): oracle vs files the reference wrote, host logic through a fake
backend (no GPU), and the CUDA path through the C ABI (`-m gpu`)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import PORTABLE_ENV, ROOT, stage_e2e

BI_CASES = ["cross", "tiny"]

_ORACLE_CHILD = r"""
import sys
sys.path.insert(0, %r)
from oracle import bi_oracle
got = bi_oracle.block_info_csv("bi.conf")
exp = open("expected.blockinfo.csv").read()
assert got == exp, "oracle CSV differs from the reference's"
print("OK", got.count("\n") - 1)
"""

# the product driver; FAKE=1 injects a numpy stand-in for the CUDA library (host-logic test without a GPU)
_PRODUCT_CHILD = r"""
import contextlib, io, os, sys
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
from wgdi_b200 import _bi, base
from wgdi_b200.block_info import block_info
if os.environ.get("FAKE") == "1":
    from test_block_info import fake_block_stats
    _bi.block_stats = fake_block_stats
opts = [(k, v) for k, v in base.load_conf("bi.conf", "blockinfo") if k != "savefile"] + [("savefile", "got.csv")]
if os.environ.get("INGEST"):
    opts.append(("ingest", os.environ["INGEST"]))
if os.environ.get("BI_REVERSE"):
    opts.append(("blast_reverse", os.environ["BI_REVERSE"]))
with contextlib.redirect_stdout(io.StringIO()):
    b = block_info(opts)
    b.run()
got, exp = open("got.csv").read(), open("expected.blockinfo.csv").read()
assert got == exp, "product CSV differs from the reference's"
print("OK", got.count("\n") - 1, b.timing)
"""


def fake_block_stats(device, hit_off, hit_g2, hit_rank, hit_order2, blk_off, mem_g1, mem_g2, mem_order2, repeat_number):
    """numpy/Python stand-in for icl_bi_block_stats with the semantics include/icl_blockinfo.h documents."""
    B = len(blk_off) - 1
    homo_sum = np.zeros((B, 5), np.int32)
    homo_cnt = np.zeros(B, np.int32)
    rows = np.zeros(B, np.int64)
    uniq = np.zeros(B, np.int32)
    for b in range(B):
        a0, a1 = int(blk_off[b]), int(blk_off[b + 1])
        subj = set(int(x) for x in mem_g2[a0:a1])
        seen = set()
        for j in range(a0, a1):
            g1, g2 = int(mem_g1[j]), int(mem_g2[j])
            if g1 < 0:
                continue
            for h in range(int(hit_off[g1]), int(hit_off[g1 + 1])):
                t, r = int(hit_g2[h]), int(hit_rank[h])
                if t == g2:
                    homo_sum[b] += [1 if r < i else 0 if r < i + 5 else -1 for i in range(1, 6)]
                    homo_cnt[b] += 1
                if g2 >= 0 and t in subj and 0 < abs(int(hit_order2[h]) - int(mem_order2[j])) <= repeat_number:
                    rows[b] += 1
                    seen.add(g1)
        uniq[b] = len(seen)
    return homo_sum, homo_cnt, rows, uniq


def _run_child(code, case, env_extra=None):
    with tempfile.TemporaryDirectory() as d:
        stage_e2e(case, d)
        env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=PORTABLE_ENV, **(env_extra or {}))
        r = subprocess.run([sys.executable, "-c", code], cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "OK" in r.stdout
        return r.stdout


@pytest.mark.parametrize("case", BI_CASES)
def test_bi_oracle_reproduces_reference_csv(case):
    _run_child(_ORACLE_CHILD % ROOT, case)


@pytest.mark.parametrize("case", BI_CASES)
@pytest.mark.parametrize("ingest", ["native", "pandas"])
def test_bi_host_logic_with_fake_backend(case, ingest):
    """Parsing, ranking with the reference's tie order, frame assembly and CSV text — device statistics from a
    numpy stand-in injected by the test."""
    _run_child(_PRODUCT_CHILD % (ROOT, ROOT), case, {"FAKE": "1", "INGEST": ingest})


@pytest.mark.parametrize("ingest", ["native", "pandas"])
def test_block_info_ignores_blast_reverse(ingest):
    """`blast_reverse = true` changes nothing in the reference's -bi: the option becomes a bool (wgdi/block_info.py:16) and
    base.newblast compares it with the string 'true' (base.py:176), so the id columns are never swapped.  Checked by
    running the unmodified reference on this fixture with the key set: it writes expected.blockinfo.csv byte for byte.
    The drop-in does the same on both ingest paths."""
    _run_child(_PRODUCT_CHILD % (ROOT, ROOT), "tiny", {"FAKE": "1", "INGEST": ingest, "BI_REVERSE": "true"})


def test_bi_library_exports_declared_symbols():
    from wgdi_b200 import _bi, build
    so = build.build_bi()
    lib = ctypes.CDLL(so)
    header = open(os.path.join(ROOT, "include", "icl_blockinfo.h")).read()
    declared = set(re.findall(r"\b(icl_bi_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_bi.EXPORTS), declared ^ set(_bi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_rank_hits_is_the_reference_sort():
    """rank_hits == per-query pandas-style nargsort (numpy's unstable sort) trimmed to repeat_number, with heavy ties."""
    from wgdi_b200.engine import nargsort_desc
    from wgdi_b200 import build
    from wgdi_b200.ingest import rank_hits, rank_hits_native
    build.build_host()
    rng = np.random.default_rng(11)
    q = rng.integers(0, 400, 20000)
    for bits in (rng.integers(100, 112, len(q)).astype(np.int64), rng.integers(100, 112, len(q)).astype(np.float64) / 2):
        for rn in (3, 12, 20, 40):
            got = rank_hits(q, bits, rn)
            assert (rank_hits_native(q, bits, 400, rn) == got).all()           # libicl_host.so + numpy for the tied queries
            exp = np.full(len(q), -1, np.int64)
            for g in np.unique(q):
                rows = np.flatnonzero(q == g)
                ranked = rows[nargsort_desc(bits[rows])][:rn]
                exp[ranked] = np.arange(len(ranked))
            assert (got == exp).all()


def test_bi_without_gpu_fails_loudly():
    from conftest import have_gpu
    if have_gpu():
        pytest.skip("a GPU is present")
    from wgdi_b200 import _bi
    z = np.zeros(2, np.int64)
    with pytest.raises(_bi.BiError):
        _bi.block_stats(0, z, [], [], [], np.array([0, 1], np.int64), [0], [0], [0], 5)


# ---- CUDA path ---------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("case", BI_CASES)
def test_bi_cuda_writes_reference_csv(case):
    _run_child(_PRODUCT_CHILD % (ROOT, ROOT), case)


@pytest.mark.gpu
def test_bi_cuda_kernel_equals_stand_in_on_random_blocks():
    """icl_bi_block_stats vs the numpy stand-in on random blocks: unknown genes (-1), repeated id1 / id2 inside a
    block, empty blocks, queries without hits."""
    from wgdi_b200 import _bi
    rng = np.random.default_rng(5)
    n1, n2, rn = 500, 400, 7
    per = rng.integers(0, rn + 1, n1)
    hit_off = np.r_[0, np.cumsum(per)].astype(np.int64)
    hit_g2 = np.concatenate([rng.choice(n2, k, replace=False) for k in per]).astype(np.int32)
    hit_rank = np.concatenate([rng.permutation(k) for k in per]).astype(np.int32)
    order_of2 = rng.integers(1, 120, n2).astype(np.int32)
    sizes = rng.integers(0, 90, 300)
    sizes[[0, 17]] = 0
    blk_off = np.r_[0, np.cumsum(sizes)].astype(np.int64)
    M = int(blk_off[-1])
    mem_g1 = rng.integers(-1, 60, M).astype(np.int32)          # small range: repeated id1 inside a block
    mem_g2 = rng.integers(-1, 80, M).astype(np.int32)
    # make a fair share of members real hits
    for j in range(0, M, 3):
        g = int(rng.integers(0, n1))
        if per[g]:
            mem_g1[j] = g
            mem_g2[j] = hit_g2[hit_off[g] + int(rng.integers(0, per[g]))]
    mem_o2 = np.where(mem_g2 >= 0, order_of2[np.maximum(mem_g2, 0)], 0).astype(np.int32)
    args = (0, hit_off, hit_g2, hit_rank, order_of2[hit_g2], blk_off, mem_g1, mem_g2, mem_o2, rn)
    got = _bi.block_stats(*args)
    exp = fake_block_stats(*args)
    for a, b in zip(got, exp):
        assert (a == b).all()
    assert exp[1].sum() > 100 and exp[2].sum() > 100 and (exp[3] > 1).any()


@pytest.mark.gpu
def test_bi_cuda_on_icl_output_matches_oracle(tmp_path):
    """-icl then -bi on a larger synthetic comparison, both through the CUDA libraries; CSV identical to the oracle's."""
    import contextlib
    import io
    from oracle import bi_oracle
    from wgdi_b200 import base, synth
    from wgdi_b200.block_info import block_info
    from wgdi_b200.run_colliearity import mycollinearity
    d = str(tmp_path)
    ds = synth.make_dataset("small")
    conf = synth.write_dataset(ds, d)
    opts = dict(base.load_conf(conf, "collinearity"))
    with contextlib.redirect_stdout(io.StringIO()):
        try:
            mycollinearity(list(opts.items())).run()
        except SystemExit:
            pass
    bi = os.path.join(d, "bi.conf")
    with open(bi, "w") as f:
        f.write("[blockinfo]\n" + "".join(f"{k} = {opts[k]}\n" for k in ("blast", "gff1", "gff2", "lens1", "lens2")) +
                f"collinearity = {opts['savefile']}\nscore = 100\nevalue = 1e-5\nrepeat_number = 20\nposition = order\n"
                f"ks = none\nks_col = ks_NG86\nsavefile = {os.path.join(d, 'bi.csv')}\n")
    with contextlib.redirect_stdout(io.StringIO()):
        block_info(base.load_conf(bi, "blockinfo")).run()
    got = open(os.path.join(d, "bi.csv")).read()
    assert got.count("\n") > 50
    assert got == bi_oracle.block_info_csv(bi)


def test_native_collinearity_reader_equals_the_python_reader(tmp_path):
    """icl_host_read_collinearity (the -bi input, include/icl_host.h) against `read_collinearity_flat`, which mirrors
    wgdi/base.py:120-141: blocks, header tokens, gene rows, positions and the joined position strings; and its refusal of
    everything that is not the plain text -icl writes (the Python reader then behaves like the reference)."""
    import gzip
    import pandas as pd
    from wgdi_b200 import _host, build
    from wgdi_b200.block_info import read_collinearity_flat
    build.build_host()
    for case in BI_CASES:
        src = os.path.join(ROOT, "tests", "golden", f"e2e_{case}")
        path = os.path.join(src, "expected.collinearity.txt")
        ids, pvs, off, n1, l1, n2, l2 = read_collinearity_flat(path)
        genes1 = sorted(set(n1))[::2] + ["not_in_the_file"]          # half of the names are unknown to the gene table
        genes2 = sorted(set(n2))
        t1, t2 = _host.pack_strings(genes1), _host.pack_strings(genes2)
        nat = _host.read_collinearity(path, t1, t2)
        assert nat is not None and nat["ids"] == ids and nat["pvalues"] == pvs and np.array_equal(nat["off"], off)
        assert np.array_equal(nat["g1"], pd.Index(genes1).get_indexer(pd.Index(n1, dtype=object)))
        assert np.array_equal(nat["g2"], pd.Index(genes2).get_indexer(pd.Index(n2, dtype=object)))
        assert nat["loc1"].tolist() == [float(x) for x in l1] and nat["loc2"].tolist() == [float(x) for x in l2]
        for b in range(len(ids)):
            assert nat["block1"][b] == "_".join(str(int(float(x))) for x in l1[off[b]:off[b + 1]])
            assert nat["block2"][b] == "_".join(str(int(float(x))) for x in l2[off[b]:off[b + 1]])
    text = open(path).read()
    t = _host.pack_strings(["a"])
    variants = {"tab": text.replace(" ", "\t", 1), "crlf": text.replace("\n", "\r\n"), "float": re.sub(r"^(\S+) (\d+) ", r"\1 \2.0 ", text, count=1, flags=re.M),
                "short": text + "lonely\n", "utf8": text.replace("&", "\u00e9", 1)}
    for name, v in variants.items():
        p = tmp_path / f"{name}.txt"
        p.write_text(v, newline="")
        assert _host.read_collinearity(str(p), t, t) is None, name
    # leading / trailing blanks, members before the first header, a header without members, an empty file
    p = tmp_path / "odd.txt"
    p.write_text("x 1 y 2 1\n# Alignment 1: score=5 pvalue=0.5 N=2 a&b plus\n# Alignment 2: score=7.5 pvalue=0.25 N=2 a&b minus\n"
                 "  a 3 a 4 1  \na 5 b 6 -1\n")
    nat = _host.read_collinearity(str(p), t, t)
    ref = read_collinearity_flat(str(p))
    assert nat["ids"] == ref[0] == ["2"] and nat["pvalues"] == ref[1] == ["0.25"] and nat["off"].tolist() == [0, 2]
    assert nat["g1"].tolist() == [0, 0] and nat["g2"].tolist() == [0, -1] and nat["block2"] == ["4_6"]
    (tmp_path / "empty.txt").write_text("")
    nat = _host.read_collinearity(str(tmp_path / "empty.txt"), t, t)
    assert nat["ids"] == [] and nat["off"].tolist() == [0]


def test_frame_to_csv_is_pandas_to_csv(tmp_path):
    """The -bi CSV writer against `DataFrame.to_csv(index=None)` byte for byte: floats of every printed form, missing
    values, integer and mixed columns, the frame `block_info.assemble` builds; and the cases it hands back to pandas."""
    import pandas as pd
    from wgdi_b200.block_info import frame_to_csv
    rng = np.random.default_rng(3)
    n = 500
    floats = np.concatenate([rng.random(n - 12) * 10.0 ** rng.integers(-12, 18, n - 12),
                             [0.0, -0.0, 1.0, 12.0, 1e16, 1e-5, 0.0001, 123456789012345680.0, np.nan, np.inf, -np.inf, 1 / 3]])
    frames = {
        "bi": pd.DataFrame([[str(i), f"c{i % 3}", "7", float(i), float(i + 5), 3.0, 9.0, f"0.{i:04d}", 6 + i % 4,
                             -1 if i % 2 else 0.25, -1 if i % 2 else 0.5, *([-1, -1, -1, -1, -1] if i % 3 else [0.5, 1 / 3, 0.0, 1.0, -0.25]),
                             "1_2_3", "4_5_6", "-1_-1_-1", floats[i]] for i in range(n)],
                           columns=['id', 'chr1', 'chr2', 'start1', 'end1', 'start2', 'end2', 'pvalue', 'length', 'ks_median',
                                    'ks_average', 'homo1', 'homo2', 'homo3', 'homo4', 'homo5', 'block1', 'block2', 'ks',
                                    'tandem_ratio']),
        "all_int_minus_one": pd.DataFrame({"a": [-1] * 5, "b": [1.5] * 5, "c": ["x"] * 5}),
        "object_mix": pd.DataFrame({"a": pd.Series([1, 2.5, "s", None, float("nan")], dtype=object), "b": [1, 2, 3, 4, 5]}),
        "empty": pd.DataFrame({"a": pd.Series([], dtype=float), "b": pd.Series([], dtype=object)}),
        "needs_quotes": pd.DataFrame({"a": ["x,y", 'say "hi"', "plain"], "b": [1.0, 2.0, 3.0]}),
        "bools": pd.DataFrame({"a": [True, False], "b": [1.0, np.nan]}),
    }
    frames["bi"]["density1"] = frames["bi"]["length"] / ((frames["bi"]["end1"] - frames["bi"]["start1"]).abs() + 1)
    frames["bi"]["class1"] = 0
    for name, df in frames.items():
        a, b = tmp_path / f"{name}_ours.csv", tmp_path / f"{name}_pandas.csv"
        frame_to_csv(df, str(a))
        df.to_csv(str(b), index=None)
        assert open(a, newline="").read() == open(b, newline="").read(), name


@pytest.mark.parametrize("homo", ["none", "some", "all"])
def test_columnwise_assemble_equals_the_row_loop(homo, monkeypatch, tmp_path):
    """`block_info._assemble_columns` (no per-block Python loop) against the row loop that mirrors block_info.py:18-75: same
    frame, same column types — a homo column of -1 only stays integer, one that mixes -1 with means is float — same CSV."""
    import pandas as pd
    from wgdi_b200.block_info import block_info, frame_to_csv
    rng = np.random.default_rng({"none": 1, "some": 2, "all": 3}[homo])
    ng = 40
    gff1 = pd.DataFrame({"chr": [f"c{i % 3}" for i in range(ng)]}, index=pd.Index([f"a{i}" for i in range(ng)]))
    gff2 = pd.DataFrame({"chr": [str(i % 2 + 1) for i in range(ng)]}, index=pd.Index([f"b{i}" for i in range(ng)]))
    B = 60
    m = rng.integers(1, 9, B)
    off = np.zeros(B + 1, np.int64)
    np.cumsum(m, out=off[1:])
    M = int(off[-1])
    m1, m2 = rng.integers(-1, ng, M).astype(np.int64), rng.integers(-1, ng, M).astype(np.int64)      # some first members unknown
    loc1, loc2 = rng.integers(1, 500, M).astype(np.float64), rng.integers(1, 500, M).astype(np.float64)
    native = dict(loc1=loc1, loc2=loc2,
                  block1=["_".join(str(int(x)) for x in loc1[off[b]:off[b + 1]]) for b in range(B)],
                  block2=["_".join(str(int(x)) for x in loc2[off[b]:off[b + 1]]) for b in range(B)])
    ids, pvs = [str(b + 1) for b in range(B)], [f"0.{b:04d}" for b in range(B)]
    cnt = {"none": np.zeros(B, np.int32), "some": rng.integers(0, 3, B).astype(np.int32), "all": rng.integers(1, 5, B).astype(np.int32)}[homo]
    hs = rng.integers(-3, 9, (B, 5)).astype(np.int32)
    rows, uniq = rng.integers(0, 6, B).astype(np.int64), rng.integers(0, 4, B).astype(np.int32)
    bi = block_info.__new__(block_info)
    frames = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ICL_BI_ROWS", mode)
        frames[mode] = bi.assemble(ids, pvs, off, None, None, None, None, m1, m2, gff1, gff2, pd.Series([], dtype=float), hs, cnt,
                                   rows, uniq, native)
        frame_to_csv(frames[mode], str(tmp_path / f"{mode}.csv"))
    assert list(frames["0"].dtypes) == list(frames["1"].dtypes)
    assert frames["0"].equals(frames["1"])
    assert open(tmp_path / "0.csv").read() == open(tmp_path / "1.csv").read()
    assert 0 < len(frames["0"]) < B
