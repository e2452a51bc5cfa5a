import ctypes
import gzip
import json
import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

PORTABLE_ENV = ("AVX2 FMA3 AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL "
                "AVX512_SPR AVX512_KNL AVX512_KNM")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_clouds():
    z = np.load(os.path.join(GOLD, "clouds.npz"))
    meta = json.loads(str(z["meta"]))
    cases = []
    for m in meta:
        pre = f"c{m['case']:02d}_"
        c = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        c["options"] = [tuple(x) for x in m["options"]]
        c["case"] = m["case"]
        cases.append(c)
    return cases


@pytest.fixture(scope="session")
def clouds():
    return load_clouds()


def stage_e2e(case, dst):
    """Copy a tests/golden/e2e_<case> fixture into `dst` and unpack the BLAST table."""
    src = os.path.join(GOLD, f"e2e_{case}")
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), dst)
    with gzip.open(os.path.join(dst, "blast.m8.gz"), "rb") as g, open(os.path.join(dst, "blast.m8"), "wb") as f:
        f.write(g.read())
    return dst


@pytest.fixture()
def emu_backend(monkeypatch):
    """Host emulation of the kernel bodies (g++ build of the same source) injected as a fake backend so the
    host-side logic can be tested without a GPU.  Test infrastructure: the package itself refuses this library."""
    from wgdi_b200 import _lib, build
    so = build.build_emu(os.path.join(ROOT, "tests", "_emu"))
    lib = _lib._declare(ctypes.CDLL(so))
    assert lib.icl_backend() == b"emu"
    monkeypatch.setattr(_lib, "_LIB", lib)
    from wgdi_b200 import engine
    engine._NP_KIND.clear()                  # which numpy flavour the "device" reproduces is probed per backend
    yield lib
    engine._NP_KIND.clear()


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_engine():
    from wgdi_b200 import _lib
    from wgdi_b200 import engine
    from wgdi_b200.engine import IclEngine
    _lib._LIB = None
    engine._NP_KIND.clear()
    eng = IclEngine(0)
    assert eng.lib.icl_backend() == b"cuda"
    yield eng
    eng.close()


def check_dp_against_golden(eng_or_fn, c):
    """DP internals of one golden cloud: score / pred / depth / used in both directions."""
    for d, tag in ((0, "1"), (1, "2")):
        s, p, dp, u = eng_or_fn(d)
        assert (s == c["score" + tag]).all(), f"case {c['case']} dir {d}: score"
        assert (u.astype(bool) == (c["used" + tag] > 0)).all(), f"case {c['case']} dir {d}: used"
        assert (p == c["pred" + tag]).all(), f"case {c['case']} dir {d}: pred"
        assert (dp == c["depth" + tag]).all(), f"case {c['case']} dir {d}: depth"


def check_blocks_against_golden(blocks, pvalue, c):
    """Emitted block list (after the rounded-p filter of collinearity.py:55-56) of one golden cloud."""
    pv = blocks.p_value()
    idx = np.flatnonzero(~(pv > pvalue))
    boff = c["boff"]
    assert len(idx) == len(boff) - 1, f"case {c['case']}: block count {len(idx)} != {len(boff) - 1}"
    for j, b in enumerate(idx):
        a0, a1 = blocks.member_off[b], blocks.member_off[b + 1]
        assert (blocks.loc1[a0:a1] == c["bloc1"][boff[j]:boff[j + 1]]).all(), f"case {c['case']} block {j}: loc1"
        assert (blocks.loc2[a0:a1] == c["bloc2"][boff[j]:boff[j + 1]]).all(), f"case {c['case']} block {j}: loc2"
        assert pv[b] == c["bp"][j], f"case {c['case']} block {j}: p {pv[b]} != {c['bp'][j]}"
        assert blocks.score[b] == c["bscore"][j]


def sorted_layout(pts):
    """(pair_off, loc1, loc2, grading) of a point set in (pair, loc1, loc2) order, ties in upload order —
    the layout wgdi/run_colliearity.py:117-121,164-166 produces pair by pair."""
    o = np.lexsort((pts["loc2"], pts["loc1"], pts["pair_id"]))
    off = np.r_[0, np.cumsum(np.bincount(pts["pair_id"], minlength=pts["n_pairs"]))].astype(np.int64)
    return off, pts["loc1"][o], pts["loc2"][o], pts["grading"][o], o


def oracle_batch(pts, prm, threads=None):
    """C oracle over all pairs, ranked with numpy in THIS process (same tie order as the engine's host rank)."""
    from oracle import c_oracle, icl_oracle
    off, l1, l2, g, o = sorted_layout(pts)
    op = icl_oracle.IclParams([("grading", ",".join(map(str, prm.grading))), ("mg", f"{prm.mg1},{prm.mg2}"),
                               ("pvalue", str(prm.pvalue)), ("coverage_ratio", str(prm.coverage_ratio))])
    bt = c_oracle.Batch(off, l1, l2, g, op, threads=threads)
    bt.chain()
    bt.rank()
    bt.extract()
    return bt, (off, l1, l2, g, o)


def oracle_blocks_stable(pts, prm):
    """`oracle_blocks` with the extraction order of a STABLE sort (what pandas' nargsort(ascending=False) would give if
    numpy's argsort were stable): the rule of the engine's non-reference device-rank mode."""
    from oracle import c_oracle, icl_oracle
    off, l1, l2, g, o = sorted_layout(pts)
    op = icl_oracle.IclParams([("grading", ",".join(map(str, prm.grading))), ("mg", f"{prm.mg1},{prm.mg2}"),
                               ("pvalue", str(prm.pvalue)), ("coverage_ratio", str(prm.coverage_ratio))])
    bt = c_oracle.Batch(off, l1, l2, g, op)
    bt.chain()
    bt.order = [np.empty(bt.n, np.int32) for _ in range(2)]
    for d in range(2):
        for p in range(bt.n_pairs):
            a, b = off[p], off[p + 1]
            sc = bt.dp[d]["score"][a:b].astype(np.int64)
            m = b - a
            bt.order[d][a:b] = (m - 1 - np.argsort(sc[::-1], kind="stable"))[::-1]
    bt.extract()
    out = []
    for p in range(pts["n_pairs"]):
        for mm, a, sc, d in bt.blocks(p):
            out.append((p, 0 if d > 0 else 1, l1[mm], l2[mm], a, sc))
    return out


def oracle_blocks(pts, prm):
    """[(pair, dir, loc1[], loc2[], a, score)] of every accepted block, in emission order."""
    bt, (off, l1, l2, g, o) = oracle_batch(pts, prm)
    out = []
    for p in range(pts["n_pairs"]):
        for mm, a, sc, d in bt.blocks(p):
            out.append((p, 0 if d > 0 else 1, l1[mm], l2[mm], a, sc))
    return out


def forked_chain_points(prefix=60, arm=200, arms=3):
    """One long diagonal prefix that forks into several long arms: after the best arm is extracted together with
    the prefix, every point of the other arms is examined and rejected one by one (coverage <= 0.8) — the case the
    extraction kernel's removed-ancestor memo exists for."""
    l1, l2 = list(range(1, prefix + 1)), list(range(1, prefix + 1))
    for a in range(arms):
        r, c = prefix, prefix + 3 * a
        for i in range(arm - 20 * a):
            r += 1
            c += 1 + (a if i == 0 else 0) * 2
            l1.append(r)
            l2.append(c)
            if a and i % 7 == 3:          # side twigs
                l1.append(r)
                l2.append(c + 9)
    pts = sorted(set(zip(l1, l2)))
    l1 = np.array([x for x, _ in pts], np.int32)
    l2 = np.array([y for _, y in pts], np.int32)
    g = np.where(np.arange(len(l1)) % 5 == 0, 50, 30).astype(np.int32)
    return dict(pair_id=np.zeros(len(l1), np.int32), loc1=l1, loc2=l2, grading=g, n_pairs=1)
