"""Seeded synthetic genomes and BLAST tables of the shapes BASELINE.json names.

There is no network and the reference ships no data, so every
measurement and every parity test runs on genomes made here.  The generator follows
SURVEY.md section 8(d): an ancestral gene order; each genome is `k` fractionated
sub-genome copies (retention 0.6) with random inversions and translocations, cut into
the stated number of chromosomes; gene families with Zipf-like sizes give off-diagonal
noise hits; 8 % tandem duplicates give near-diagonal hits; bit scores are integers so
ties occur; rows are emitted per query in descending score like BLAST; sub-threshold
rows, self hits and duplicated (id1, id2) rows are mixed in so that every filter of
wgdi/base.py:178-180 fires.

Two outputs are offered: integer-coded hit arrays (fast path for the benchmark) and real
gff / lens / m8 / conf files that the unmodified reference can read.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, replace

import numpy as np


@dataclass
class SynthConfig:
    name: str
    seed: int
    n_anc: int                  # ancestral genes
    k: int                      # sub-genome copies per genome
    n_chr1: int
    n_chr2: int
    self_cmp: bool              # genome 2 is genome 1
    target_points: int          # graded points aimed at
    retention: float = 0.6
    tandem_frac: float = 0.08
    grading: tuple = (50, 30, 25)
    mg: tuple = (25, 25)
    multiple: int = 1
    pvalue: float = 1.0
    repeat_number: int | None = None      # None -> 1.2 x target/genes
    float_scores: bool = False            # a third of the bit scores get a decimal part


# The five BASELINE.json configs (seeds 20260000 + config) and small cases.
CONFIGS = {
    "tiny": SynthConfig("tiny", 20260090, n_anc=260, k=2, n_chr1=3, n_chr2=3, self_cmp=True, target_points=6_000),
    "small": SynthConfig("small", 20260091, n_anc=1100, k=2, n_chr1=4, n_chr2=4, self_cmp=True, target_points=40_000),
    "cross": SynthConfig("cross", 20260092, n_anc=700, k=2, n_chr1=4, n_chr2=3, self_cmp=False, target_points=12_000,
                         grading=(60, 35, 20), mg=(40, 30), multiple=2, pvalue=0.3, float_scores=True),
    "C1": SynthConfig("C1", 20260001, n_anc=23_150, k=2, n_chr1=5, n_chr2=5, self_cmp=True, target_points=1_000_000),
    "C2": SynthConfig("C2", 20260002, n_anc=15_430, k=3, n_chr1=19, n_chr2=19, self_cmp=True, target_points=2_000_000),
    "C3": SynthConfig("C3", 20260003, n_anc=30_860, k=2, n_chr1=12, n_chr2=12, self_cmp=False, target_points=5_000_000),
    "C4": SynthConfig("C4", 20260004, n_anc=56_580, k=3, n_chr1=21, n_chr2=21, self_cmp=True, target_points=10_000_000),
}


# BASELINE.json config 5: 8 species (genes, chromosomes) -> 28 pairwise comparisons, each swept over
# grading x mg x pvalue
C5_SPECIES = [(25_000, 7), (28_000, 9), (30_000, 10), (33_000, 12), (36_000, 13), (39_000, 16), (42_000, 17), (45_000, 19)]
C5_GRADINGS = [(50, 40, 25), (50, 30, 25)]
C5_MGS = [(25, 25), (40, 40), (50, 50)]
C5_PVALUES = [0.05, 0.2, 1.0]


def config5_comparison(ci: int) -> SynthConfig:
    """The ci-th (0..27) cross-species comparison of config 5, in itertools.combinations order of the species."""
    import itertools
    a, b = list(itertools.combinations(range(len(C5_SPECIES)), 2))[ci]
    (ga, ca), (gb, cb) = C5_SPECIES[a], C5_SPECIES[b]
    genes = (ga + gb) // 2
    return SynthConfig(f"c5_{a}_{b}", 20260500 + ci, n_anc=int(genes / 2 / 0.6), k=2, n_chr1=ca, n_chr2=cb,
                       self_cmp=False, target_points=int(genes * 60))


@dataclass
class Genome:
    prefix: str
    anc: np.ndarray        # ancestral gene id per gene (genome order)
    chr_idx: np.ndarray    # 0-based chromosome index per gene
    order: np.ndarray      # 1-based gene order inside its chromosome
    strand: np.ndarray     # '+' / '-'
    chr_names: list

    @property
    def n(self):
        return len(self.anc)

    def ids(self):
        return np.array([f"{self.prefix}{self.chr_names[c]}g{o:05d}" for c, o in zip(self.chr_idx, self.order)],
                        dtype=object)


def _make_genome(rng, cfg: SynthConfig, n_chr: int, prefix: str, chr_names=None) -> Genome:
    subs = []
    for _ in range(cfg.k):
        order = np.flatnonzero(rng.random(cfg.n_anc) < cfg.retention)
        n_events = max(2, len(order) // 1500)
        for _ in range(n_events):                         # inversions
            a, b = np.sort(rng.integers(0, len(order), 2))
            order[a:b] = order[a:b][::-1].copy()
        for _ in range(n_events // 2):                    # translocations
            a, b = np.sort(rng.integers(0, len(order), 2))
            seg, rest = order[a:b], np.r_[order[:a], order[b:]]
            at = int(rng.integers(0, len(rest) + 1))
            order = np.r_[rest[:at], seg, rest[at:]]
        subs.append(order)
    anc = np.concatenate(subs)
    anc = np.repeat(anc, 1 + (rng.random(len(anc)) < cfg.tandem_frac))       # tandem duplicates
    n = len(anc)
    # unequal chromosome sizes (+-25 %)
    w = rng.uniform(0.75, 1.25, n_chr)
    bounds = np.r_[0, np.round(np.cumsum(w) / w.sum() * n).astype(np.int64)]
    bounds[-1] = n
    chr_idx = np.searchsorted(bounds, np.arange(n), side="right") - 1
    order = np.arange(n) - bounds[chr_idx] + 1
    strand = np.where(rng.random(n) < 0.5, "+", "-")
    names = chr_names or [str(i + 1) for i in range(n_chr)]
    return Genome(prefix, anc.astype(np.int64), chr_idx.astype(np.int64), order.astype(np.int64), strand, names)


@dataclass
class Dataset:
    cfg: SynthConfig
    g1: Genome
    g2: Genome
    q: np.ndarray          # query gene index (into g1) per BLAST row, file order
    t: np.ndarray          # subject gene index (into g2)
    bits: np.ndarray       # bit score (int64, or float64 when cfg.float_scores)
    evalue: np.ndarray     # float64
    repeat_number: int

    def options(self, paths=None, savefile="out.collinearity.txt", process=8):
        """`[collinearity]` options as configparser would return them (all strings)."""
        p = paths or {}
        c = self.cfg
        return [
            ("gff1", p.get("gff1", "")), ("gff2", p.get("gff2", "")),
            ("lens1", p.get("lens1", "")), ("lens2", p.get("lens2", "")),
            ("blast", p.get("blast", "")), ("blast_reverse", "false"),
            ("comparison", "genomes"), ("multiple", str(c.multiple)), ("process", str(process)),
            ("evalue", "1e-5"), ("score", "100"),
            ("grading", ",".join(str(x) for x in c.grading)), ("mg", ",".join(str(x) for x in c.mg)),
            ("pvalue", repr(float(c.pvalue)) if c.pvalue != 1 else "1"),
            ("repeat_number", str(self.repeat_number)), ("positon", "order"),
            ("savefile", savefile),
        ]


def make_dataset(cfg: SynthConfig | str, scale: float = 1.0) -> Dataset:
    """Build genomes and the raw BLAST table.  `scale` < 1 shrinks genes and hits together
    (used to take bounded samples of a named workload)."""
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    if scale != 1.0:
        cfg = replace(cfg, n_anc=max(60, int(cfg.n_anc * scale)), target_points=max(500, int(cfg.target_points * scale)))
    rng = np.random.default_rng(cfg.seed)
    g1 = _make_genome(rng, cfg, cfg.n_chr1, "a")
    g2 = g1 if cfg.self_cmp else _make_genome(rng, cfg, cfg.n_chr2, "b")

    per_gene = cfg.target_points / g1.n
    repeat_number = cfg.repeat_number or max(6, int(np.ceil(1.2 * per_gene)))

    # --- homeolog hits: every pair of genes descending from the same ancestral gene ---------
    o2 = np.argsort(g2.anc, kind="stable")
    start2 = np.searchsorted(g2.anc[o2], np.arange(cfg.n_anc + 1))
    cnt = (start2[1:] - start2[:-1])[g1.anc]
    hq = np.repeat(np.arange(g1.n), cnt)
    within = np.arange(len(hq)) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    ht = o2[start2[g1.anc][hq] + within]
    hbits = np.clip(rng.normal(600, 150, len(hq)), 100, None)
    if cfg.self_cmp:                                     # the self hit is the best hit, as in real BLAST
        hbits = np.where(hq == ht, 2000.0, hbits)

    # --- family / noise hits ---------------------------------------------------------------
    fam_of_anc = np.minimum(rng.zipf(1.6, cfg.n_anc), 400) * 1000 + rng.integers(0, 60, cfg.n_anc)
    fam2 = fam_of_anc[g2.anc]
    f2 = np.argsort(fam2, kind="stable")
    fam_sorted = fam2[f2]
    lam = per_gene * 1.12
    nn = rng.poisson(lam, g1.n)
    nq = np.repeat(np.arange(g1.n), nn)
    in_family = rng.random(len(nq)) < 0.35
    fq = fam_of_anc[g1.anc][nq]
    lo = np.searchsorted(fam_sorted, fq, side="left")
    hi = np.searchsorted(fam_sorted, fq, side="right")
    pick = lo + (rng.random(len(nq)) * np.maximum(hi - lo, 1)).astype(np.int64)
    nt = np.where(in_family & (hi > lo), f2[np.minimum(pick, g2.n - 1)], rng.integers(0, g2.n, len(nq)))
    nbits = 100 + rng.exponential(80, len(nq))

    q = np.r_[hq, nq]
    t = np.r_[ht, nt]
    bits = np.r_[hbits, nbits]
    # 5 % sub-threshold rows: half by score, half by evalue
    ev = np.power(10.0, -np.minimum(bits, 1200) / 4.0)
    sub = rng.random(len(q))
    bits = np.where(sub < 0.025, rng.uniform(40, 99.4, len(q)), bits)
    ev = np.where((sub >= 0.025) & (sub < 0.05), rng.uniform(2e-5, 1e-2, len(q)), ev)
    ev = np.where((sub >= 0.05) & (sub < 0.052), 1e-5, ev)      # exactly the threshold: strict < drops it
    if cfg.float_scores:
        frac = np.where(rng.random(len(q)) < 0.33, 0.5, 0.0)
        bits = np.round(bits) + frac
    else:
        bits = np.round(bits).astype(np.int64)
    # duplicated (id1, id2) rows with a lower score, placed later in the query's block
    dup = rng.random(len(q)) < 0.004
    q = np.r_[q, q[dup]]
    t = np.r_[t, t[dup]]
    bits = np.r_[bits, np.maximum(bits[dup] - 7, 30)]
    ev = np.r_[ev, ev[dup]]
    # BLAST order: grouped by query, best hit first (stable: ties keep generation order)
    o = np.lexsort((-bits, q))
    return Dataset(cfg, g1, g2, q[o], t[o], bits[o], ev[o], repeat_number)


def write_dataset(ds: Dataset, outdir: str, savefile: str | None = None, process: int = 8) -> str:
    """Write gff/lens/m8 files + a conf the unmodified reference can run; returns the conf path."""
    os.makedirs(outdir, exist_ok=True)
    paths = {}
    genomes = [("1", ds.g1)] + ([("2", ds.g2)] if not ds.cfg.self_cmp else [])
    idcache = {}
    for tag, g in genomes:
        ids = g.ids()
        idcache[tag] = ids
        gff = os.path.join(outdir, f"genome{tag}.gff")
        with open(gff, "w") as f:
            for i in range(g.n):
                st = 1000 * int(g.order[i])
                f.write(f"{g.chr_names[g.chr_idx[i]]}\t{ids[i]}\t{st}\t{st + 800}\t{g.strand[i]}\t{int(g.order[i])}\t{ids[i]}.1\n")
        lens = os.path.join(outdir, f"genome{tag}.lens")
        with open(lens, "w") as f:
            for c, name in enumerate(g.chr_names):
                ng = int((g.chr_idx == c).sum())
                f.write(f"{name}\t{ng * 1000 + 800}\t{ng}\n")
        paths[f"gff{tag}"], paths[f"lens{tag}"] = gff, lens
    if ds.cfg.self_cmp:
        paths["gff2"], paths["lens2"] = paths["gff1"], paths["lens1"]
        idcache["2"] = idcache["1"]
    blast = os.path.join(outdir, "blast.m8")
    i1, i2 = idcache["1"][ds.q], idcache["2"][ds.t]
    with open(blast, "w") as f:
        for a, b, s, e in zip(i1, i2, ds.bits, ds.evalue):
            sc = f"{s:.1f}" if ds.cfg.float_scores and s != int(s) else f"{int(s)}"
            f.write(f"{a}\t{b}\t87.50\t200\t20\t2\t1\t200\t1\t200\t{e:.2e}\t{sc}\n")
    paths["blast"] = blast
    conf = os.path.join(outdir, "icl.conf")
    savefile = savefile or os.path.join(outdir, "out.collinearity.txt")
    with open(conf, "w") as f:
        f.write("[collinearity]\n")
        for k, v in ds.options(paths, savefile, process):
            f.write(f"{k} = {v}\n")
    return conf


def graded_points(ds: Dataset, grading=None) -> dict:
    """Host ingest of the integer-coded BLAST table: the filters of wgdi/base.py:178-180, the grading of
    wgdi/run_colliearity.py:59-86 and the (chr1, chr2) grouping of :117-121, vectorised — the same code
    path `mycollinearity.load_points` takes after parsing the files.
    `grading`: grade with these values instead of the config's, straight through `grade_hits` (what `-icl` with that
    conf value does; the independent check of the class-table path of a grading sweep)."""
    import pandas as pd
    from .ingest import grade_hits, pair_ids
    c = ds.cfg
    keep = (ds.bits >= 100) & (ds.evalue < 1e-5)
    if c.self_cmp:
        keep &= ds.q != ds.t
    rows = np.flatnonzero(keep)
    q, t = ds.q[rows], ds.t[rows]
    first = ~pd.Series(q.astype(np.int64) * ds.g2.n + t).duplicated(keep="first").to_numpy()
    rows, q, t = rows[first], q[first], t[first]
    # rank class of every hit (0 red, 1 blue, 2 rest; run_colliearity.py:64-86) — the grading is a table lookup on it, so a
    # sweep over `grading` keeps the classes and swaps the table (icl_set_grading_table)
    cls = grade_hits(q, ds.bits[rows], c.multiple, ds.repeat_number, [1, 2, 3])
    ok = cls > 0
    q, t, cls = q[ok], t[ok], cls[ok] - 1
    grade = np.asarray(c.grading, np.int64)[cls]
    if grading is not None:
        direct = grade_hits(ds.q[rows], ds.bits[rows], c.multiple, ds.repeat_number, list(grading))
        grade = direct[ok]
    names1 = sorted(ds.g1.chr_names)
    names2 = sorted(ds.g2.chr_names)
    code1 = np.array([names1.index(x) for x in ds.g1.chr_names])[ds.g1.chr_idx[q]]
    code2 = np.array([names2.index(x) for x in ds.g2.chr_names])[ds.g2.chr_idx[t]]
    pid, names = pair_ids(names1, names2, code1, code2)
    return dict(pair_id=pid, loc1=ds.g1.order[q].astype(np.int32), loc2=ds.g2.order[t].astype(np.int32),
                grading=grade.astype(np.int32), grade_class=cls.astype(np.int32), n_pairs=len(names), pair_names=names)
