"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run once in the build container (the only place /root/reference exists):

    python tests/make_golden.py            # inner-seam clouds + end-to-end files

Nothing in the test-suite imports the reference; the tests replay the fixtures written
here.  The reference ships no tests or golden vectors of its own, so
these reference-generated outputs are the pin for the oracle and for the CUDA path.

Two kinds of fixture:

* `clouds.npz` — inner seam `wgdi.collinearity.collinearity(options, points).run()`
  (wgdi/collinearity.py:42-70) on random point clouds.  Besides the returned blocks we
  store what the reference left in its frame after `run()` (score1/2, path1/2,
  usedtimes1/2) and the descending-score permutation pandas produced in the generating
  process, so DP and extraction can be checked independently of the CPU-specific tie
  order of numpy's unstable argsort (SURVEY.md hazard H1).
* `e2e_<case>/` — outer seam `mycollinearity(options).run()` (wgdi/run_colliearity.py:88-156)
  on small synthetic genomes: input files + the file the reference wrote.  These are
  generated in a child process with numpy's SIMD sort kernels disabled
  (NPY_DISABLE_CPU_FEATURES), which selects numpy's scalar introsort and makes the tie
  order reproducible on any x86 machine; the tests replay them in a child process with
  the same setting and check the argsort fingerprint first.
"""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
REFERENCE = "/root/reference"

PORTABLE_ENV = ("AVX2 FMA3 AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL "
                "AVX512_SPR AVX512_KNL AVX512_KNM")


def argsort_fingerprint():
    import numpy as np
    a = (np.arange(1000) * 7919 % 5).astype(np.int64)
    return hashlib.md5(np.argsort(a, kind="quicksort").tobytes()).hexdigest()


def import_reference():
    """wgdi/base.py:6-10 imports matplotlib and Bio, unused on the -icl path: stub them."""
    for m in ("matplotlib", "matplotlib.patches", "matplotlib.pyplot", "Bio", "Bio.SeqIO"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.modules["Bio"].SeqIO = sys.modules["Bio.SeqIO"]
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import wgdi.base as base
    import wgdi.collinearity as refcol
    from wgdi.run_colliearity import mycollinearity
    return base, refcol, mycollinearity


# ----------------------------------------------------------------------------------------
# inner seam
# ----------------------------------------------------------------------------------------

def make_cloud(rng, n, grid1, grid2, diag_frac, dup=False):
    import numpy as np
    pts = set()
    n_diag = int(n * diag_frac)
    while len(pts) < n_diag:
        r0, c0 = int(rng.integers(0, grid1)), int(rng.integers(0, grid2))
        length = int(rng.integers(4, 40))
        sign = 1 if rng.random() < 0.5 else -1
        r, c = r0, c0
        for _ in range(length):
            if 0 <= r < grid1 and 0 <= c < grid2 and rng.random() < 0.8:
                pts.add((r + 1, c + 1))
            r += int(rng.integers(1, 4))
            c += sign * int(rng.integers(1, 4))
    while len(pts) < n:
        pts.add((int(rng.integers(1, grid1 + 1)), int(rng.integers(1, grid2 + 1))))
    pts = sorted(pts)[:n]
    loc1 = np.array([p[0] for p in pts], np.int64)
    loc2 = np.array([p[1] for p in pts], np.int64)
    if dup and len(loc1) > 10:                   # H22: equal (loc1, loc2) keeps file order
        loc1 = np.r_[loc1, loc1[5:8]]
        loc2 = np.r_[loc2, loc2[5:8]]
        o = np.lexsort((loc2, loc1))
        loc1, loc2 = loc1[o], loc2[o]
    return loc1, loc2


def gen_clouds():
    import numpy as np
    import pandas as pd
    _, refcol, _ = import_reference()
    rng = np.random.default_rng(20260101)
    cases = []
    spec = []
    mgs = [(25, 25), (40, 40), (10, 15), (5, 30), (50, 50), (33, 64)]
    gradings = [(50, 30, 25), (50, 40, 25), (60, 35, 20), (100, 60, 30)]
    for ci in range(44):
        n = int(rng.choice([3, 8, 60, 120, 200, 300, 400]))
        if ci == 0:
            n = 0
        if ci == 1:
            n = 1
        grid1 = int(rng.choice([30, 60, 120, 150]))
        grid2 = int(rng.choice([30, 60, 120, 150]))
        n = min(n, grid1 * grid2 // 2)
        spec.append(dict(n=n, grid1=grid1, grid2=grid2, diag=float(rng.choice([0.0, 0.3, 0.7])),
                         mg=mgs[ci % len(mgs)], grading=gradings[(ci // 2) % len(gradings)],
                         pvalue=float(rng.choice([1, 0.2, 0.05])), cov=0.8 if ci % 7 else 0.6, dup=(ci % 11 == 5)))
    out = {}
    meta = []
    for ci, s in enumerate(spec):
        loc1, loc2 = make_cloud(rng, s["n"], s["grid1"], s["grid2"], s["diag"], s["dup"])
        n = len(loc1)
        grading = rng.choice(np.array(s["grading"], np.int64), size=n, p=[0.2, 0.3, 0.5])
        labels = rng.permutation(np.arange(1000, 1000 + n))          # BLAST row labels: unique, unordered
        points = pd.DataFrame({"loc1": loc1, "loc2": loc2, "grading": grading}, index=labels)
        options = [("grading", ",".join(map(str, s["grading"]))), ("mg", f"{s['mg'][0]},{s['mg'][1]}"),
                   ("pvalue", str(s["pvalue"])), ("coverage_ratio", str(s["cov"]))]
        ref = refcol.collinearity(options, points.copy())
        data = ref.run() if n > 0 else []
        pos = {lab: i for i, lab in enumerate(labels)}
        if n > 0:
            fr = ref.points
            sc1 = fr["score1"].to_numpy(np.int64)
            sc2 = fr["score2"].to_numpy(np.int64)
            u1 = fr["usedtimes1"].to_numpy(np.int64)
            u2 = fr["usedtimes2"].to_numpy(np.int64)
            pr1 = np.array([pos[p[-2]] if len(p) > 1 else -1 for p in fr["path1"]], np.int64)
            pr2 = np.array([pos[p[-2]] if len(p) > 1 else -1 for p in fr["path2"]], np.int64)
            d1 = np.array([len(p) for p in fr["path1"]], np.int64)
            d2 = np.array([len(p) for p in fr["path2"]], np.int64)
            # same call, same process -> same permutation the reference used (collinearity.py:49,60)
            o1 = fr[["score1"]].sort_values(by=["score1"], ascending=False).index.map(pos).to_numpy(np.int64)
            o2 = fr[["score2"]].sort_values(by=["score2"], ascending=False).index.map(pos).to_numpy(np.int64)
        else:
            sc1 = sc2 = u1 = u2 = pr1 = pr2 = d1 = d2 = o1 = o2 = np.zeros(0, np.int64)
        boff = [0]
        b1, b2, bp, bs = [], [], [], []
        for block, p, score in data:
            b1.append(block["loc1"].to_numpy(np.int64))
            b2.append(block["loc2"].to_numpy(np.int64))
            bp.append(float(p))
            bs.append(int(score))
            boff.append(boff[-1] + len(block))
        pre = f"c{ci:02d}_"
        out[pre + "loc1"], out[pre + "loc2"], out[pre + "grading"] = loc1, loc2, grading
        out[pre + "score1"], out[pre + "score2"] = sc1, sc2
        out[pre + "used1"], out[pre + "used2"] = u1, u2
        out[pre + "pred1"], out[pre + "pred2"] = pr1, pr2
        out[pre + "depth1"], out[pre + "depth2"] = d1, d2
        out[pre + "order1"], out[pre + "order2"] = o1, o2
        out[pre + "boff"] = np.asarray(boff, np.int64)
        out[pre + "bloc1"] = np.concatenate(b1) if b1 else np.zeros(0, np.int64)
        out[pre + "bloc2"] = np.concatenate(b2) if b2 else np.zeros(0, np.int64)
        out[pre + "bp"] = np.asarray(bp, np.float64)
        out[pre + "bscore"] = np.asarray(bs, np.int64)
        meta.append(dict(case=ci, options=options, n=n, blocks=len(bp)))
        print(f"cloud {ci}: n={n} mg={s['mg']} blocks={len(bp)}", flush=True)
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "clouds.npz"), **out)


# ----------------------------------------------------------------------------------------
# outer seam (runs in a child with the portable numpy sort)
# ----------------------------------------------------------------------------------------

E2E_CASES = {
    # name -> (synth config, overrides written into the conf, renames)
    "tiny": dict(cfg="tiny"),
    "cross": dict(cfg="cross"),
    "tiny_rev": dict(cfg="cross", blast_reverse=True, chr_names1=["10", "B2", "1", "2"]),
    # `comparison = chromosomes` (wgdi/run_colliearity.py:42-57).  The reference's code predates pandas 3: its
    # groupby(...).apply(f) at :55 relies on f receiving the grouping columns, which pandas 3 removed (KeyError
    # 'chr1' at :118).  The reference is still run UNMODIFIED; the generating process restores that
    # one pandas behaviour (`pandas2_groupby_apply_shim`) so that the mode can be pinned at all.
    "cross_chrom": dict(cfg="cross", comparison="chromosomes", pandas2_shim=True),
}


def pandas2_groupby_apply_shim():
    """pandas < 3 semantics of DataFrameGroupBy.apply for the one call the reference makes
    (wgdi/run_colliearity.py:55): the function receives every group WITH its grouping columns, groups come in
    sorted key order, results are concatenated (the caller resets the index right away)."""
    import pandas as pd
    from pandas.core.groupby.generic import DataFrameGroupBy

    def apply_with_groups(self, func, *a, **k):
        pieces = [func(g, *a, **k) for _, g in self]
        return pd.concat(pieces) if pieces else self.obj.iloc[:0]
    DataFrameGroupBy.apply = apply_with_groups


def gen_e2e_child(case):
    sys.path.insert(0, ROOT)
    import numpy as np
    from wgdi_b200 import synth
    base, _, mycollinearity = import_reference()
    spec = E2E_CASES[case]
    ds = synth.make_dataset(spec["cfg"])
    if "chr_names1" in spec:
        ds.g1.chr_names = spec["chr_names1"]
    d = os.path.join(GOLD, f"e2e_{case}")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    cwd = os.getcwd()
    os.chdir(d)
    try:
        if spec.get("blast_reverse"):
            ds_q, ds_t = ds.q, ds.t
        conf = synth.write_dataset(ds, ".", savefile="expected.collinearity.txt", process=4)
        if spec.get("blast_reverse"):
            # swap the two id columns in the file and ask the reader to swap them back (base.py:176-177)
            lines = open("blast.m8").read().splitlines()
            with open("blast.m8", "w") as f:
                for ln in lines:
                    c = ln.split("\t")
                    c[0], c[1] = c[1], c[0]
                    f.write("\t".join(c) + "\n")
            txt = open(conf).read().replace("blast_reverse = false", "blast_reverse = true")
            open(conf, "w").write(txt)
        if spec.get("comparison"):
            txt = open(conf).read().replace("comparison = genomes", f"comparison = {spec['comparison']}")
            open(conf, "w").write(txt)
        if spec.get("pandas2_shim"):
            pandas2_groupby_apply_shim()
        try:
            mycollinearity(base.load_conf(conf, "collinearity")).run()
        except SystemExit:
            pass
        with open("blast.m8", "rb") as f, gzip.GzipFile("blast.m8.gz", "wb", mtime=0) as g:
            g.write(f.read())
        os.remove("blast.m8")
        json.dump({"argsort_fingerprint": argsort_fingerprint(), "numpy": np.__version__,
                   "NPY_DISABLE_CPU_FEATURES": PORTABLE_ENV, "pandas2_groupby_apply_shim": bool(spec.get("pandas2_shim"))},
                  open("meta.json", "w"), indent=1)
    finally:
        os.chdir(cwd)
    print(case, "blocks:", sum(1 for l in open(os.path.join(d, "expected.collinearity.txt")) if l.startswith("#")))


# ----------------------------------------------------------------------------------------
# `-bi` (block information) on the collinearity files the reference wrote above
# ----------------------------------------------------------------------------------------

BI_CASES = {"cross": dict(ks=False, repeat_number=20), "tiny": dict(ks=True, repeat_number=12)}


def gen_bi_child(case):
    """wgdi.block_info.block_info(options).run() (wgdi/block_info.py:136-165), unmodified, on e2e_<case>."""
    import numpy as np
    base, _, _ = import_reference()
    from wgdi.block_info import block_info
    spec = BI_CASES[case]
    d = os.path.join(GOLD, f"e2e_{case}")
    cwd = os.getcwd()
    os.chdir(d)
    try:
        with gzip.open("blast.m8.gz", "rb") as g, open("blast.m8", "wb") as f:
            f.write(g.read())
        icl = dict(base.load_conf("icl.conf", "collinearity"))
        ks = "none"
        if spec["ks"]:
            # synthetic Ks table over the block pairs: some given reversed, some missing, negative values (dropped
            # by read_ks), duplicated keys (first wins)
            rng = np.random.default_rng(42)
            pairs = [ln.split()[0:3:2] for ln in open("expected.collinearity.txt") if not ln.startswith("#")]
            with open("ks.tsv", "w") as f:
                f.write("id1\tid2\tka_NG86\tks_NG86\tka_YN00\tks_YN00\n")
                for a, b in pairs:
                    u = rng.random()
                    if u < 0.15:
                        continue
                    if u < 0.35:
                        a, b = b, a
                    v = -1.0 if u > 0.93 else round(float(rng.gamma(2.0, 0.4)), 4)
                    f.write(f"{a}\t{b}\t{round(v / 3, 4)}\t{v}\t0.1\t{round(abs(v) * 1.1, 4)}\n")
                    if u > 0.6 and u < 0.65:
                        f.write(f"{a}\t{b}\t0.5\t9.9999\t0.1\t9.9\n")
            ks = "ks.tsv"
        with open("bi.conf", "w") as f:
            f.write("[blockinfo]\n"
                    f"blast = blast.m8\ngff1 = {icl['gff1']}\ngff2 = {icl['gff2']}\nlens1 = {icl['lens1']}\nlens2 = {icl['lens2']}\n"
                    f"collinearity = expected.collinearity.txt\nscore = 100\nevalue = 1e-5\nrepeat_number = {spec['repeat_number']}\n"
                    f"position = order\nks = {ks}\nks_col = ks_NG86\nsavefile = expected.blockinfo.csv\n")
        block_info(base.load_conf("bi.conf", "blockinfo")).run()
        os.remove("blast.m8")
    finally:
        os.chdir(cwd)
    print(case, "blockinfo rows:", sum(1 for _ in open(os.path.join(d, "expected.blockinfo.csv"))) - 1)


COLOR_PARAMS = [(1, 20), (2, 12), (3, 5), (1, 3), (0, 8), (4, 40)]      # (multiple, repeat_number)


def gen_colors_child(case):
    """dotplot.pair_positon (wgdi/dotplot.py:31-56), unmodified, on the BLAST table of e2e_<case>: for several (multiple,
    repeat_number) the rows it keeps and their colours -> e2e_<case>/expected.colors.npz."""
    import numpy as np
    base, _, _ = import_reference()
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from wgdi.dotplot import dotplot
    d = os.path.join(GOLD, f"e2e_{case}")
    cwd = os.getcwd()
    os.chdir(d)
    try:
        icl = dict(base.load_conf("icl.conf", "collinearity"))
        lens1, lens2 = base.newlens(icl["lens1"], "order"), base.newlens(icl["lens2"], "order")
        gff1, gff2 = base.newgff(icl["gff1"]), base.newgff(icl["gff2"])
        # dotplot.py:97-98 calls base.gene_location, which raises under pandas >= 3 (it assigns floats into a column created
        # as '' — TypeError: Invalid value for dtype 'str'); pair_positon only maps the column, so the fixture gives it the
        # lens filter of base.py:222 and a plain numeric 'loc'
        gff1 = gff1[gff1["chr"].isin(lens1.index)].copy()
        gff2 = gff2[gff2["chr"].isin(lens2.index)].copy()
        gff1["loc"] = gff1["order"].astype(float)
        gff2["loc"] = gff2["order"].astype(float)
        out = {}
        for reverse in ("false", "true"):
            for mult, rn in COLOR_PARAMS:
                blast = base.newblast("blast.m8.gz", 100, 1e-5, gff1, gff2, reverse)  # :108-109
                df = dotplot.pair_positon(None, blast, gff1, gff2, mult, rn)
                code = df["color"].map({"red": 1, "blue": 2, "gray": 3}).to_numpy().astype(np.int8)
                out[f"{reverse}_{mult}_{rn}_row"] = df.index.to_numpy().astype(np.int64)
                out[f"{reverse}_{mult}_{rn}_color"] = code
        np.savez_compressed("expected.colors.npz", **out)
    finally:
        os.chdir(cwd)
    print(case, "colour sets:", len(out) // 2)


def gen_colors():
    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=PORTABLE_ENV)
    for case in ("tiny", "cross"):
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--colors-child", case], env=env)


def gen_bi():
    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=PORTABLE_ENV)
    for case in BI_CASES:
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--bi-child", case], env=env)


def gen_e2e(only=None):
    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=PORTABLE_ENV)
    for case in E2E_CASES:
        if only and case != only:
            continue
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--e2e-child", case], env=env)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--e2e-child":
        gen_e2e_child(sys.argv[2])
    elif len(sys.argv) > 2 and sys.argv[1] == "--bi-child":
        gen_bi_child(sys.argv[2])
    elif len(sys.argv) > 2 and sys.argv[1] == "--colors-child":
        gen_colors_child(sys.argv[2])
    else:
        what = sys.argv[1] if len(sys.argv) > 1 else "all"
        if what in ("all", "clouds"):
            gen_clouds()
        if what in ("all", "e2e"):
            gen_e2e()
        if what in E2E_CASES:
            gen_e2e(only=what)
        if what in ("all", "bi"):
            gen_bi()
        if what in ("all", "colors"):
            gen_colors()
