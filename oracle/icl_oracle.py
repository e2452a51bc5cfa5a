"""CPU oracle for the `wgdi -icl` hot path (TEST INFRASTRUCTURE, not product code).

This is a plain numpy/Python restatement of the algorithm the reference runs for
`wgdi -icl` (wgdi/run_colliearity.py driving wgdi/collinearity.py).  It exists so the
CUDA path can be checked bit for bit.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import it; the product package
`wgdi_b200` never does.

Parity pin: the reference ships no tests or golden vectors, so this
restatement is pinned against the *reference itself*, executed in the build container by
`tests/make_golden.py` (which imports /root/reference/wgdi with matplotlib/Bio stubbed).
The resulting fixtures live in `tests/golden/` and `tests/test_oracle_golden.py` replays
them on every run.

Every function cites the reference lines it restates (paths relative to the reference
checkout, e.g. wgdi/collinearity.py:72-115).

Third-party arithmetic the reference relies on and that is therefore reproduced by
calling the same library, not by re-implementing it: numpy.argsort(kind='quicksort')
reached through pandas.core.sorting.nargsort (unstable, CPU-dependent tie order),
numpy float64 rounding (`round(np.float64, 4)`).
"""
from __future__ import annotations

import configparser

import numpy as np

# --------------------------------------------------------------------------------------
# Parameters
# --------------------------------------------------------------------------------------


class IclParams:
    """Typed view of the `[collinearity]` options the DP engine reads.

    Restates wgdi/collinearity.py:6-28 (defaults gap_penalty=-1, over_gap=3, mg 40/40,
    pvalue 1, coverage_ratio 0.8, grading 50,40,25; every option key is first set as a
    string attribute and then grading/mg/pvalue/coverage_ratio are parsed).
    """

    def __init__(self, options=()):
        opts = {str(k): v for k, v in options}
        self.grading = [int(k) for k in str(opts["grading"]).split(",")] if "grading" in opts else [50, 40, 25]
        self.mg1, self.mg2 = [int(k) for k in str(opts["mg"]).split(",")] if "mg" in opts else [40, 40]
        self.pvalue = float(opts.get("pvalue", 1))
        self.coverage_ratio = float(opts.get("coverage_ratio", 0.8))
        self.over_gap = 3          # inner over_gap, collinearity.py:13 (not a usable conf key)
        self.gap_penalty = -1      # collinearity.py:8


# --------------------------------------------------------------------------------------
# numpy call sites that decide tie order (hazards H1/H2)
# --------------------------------------------------------------------------------------


def nargsort_desc(a: np.ndarray) -> np.ndarray:
    """pandas.core.sorting.nargsort(a, kind='quicksort', ascending=False) for a NaN-free array.

    This is what `DataFrame.sort_values(by=[col], ascending=False)` executes for a single
    key (wgdi/collinearity.py:49,60 and wgdi/run_colliearity.py:67): reverse, unstable
    argsort, reverse.  The dtype and the element order of `a` both influence the tie
    permutation, so callers must pass exactly the column the reference sorts.
    """
    a = np.asarray(a)
    idx = np.arange(len(a))
    return idx[::-1][a[::-1].argsort(kind="quicksort")][::-1]


# --------------------------------------------------------------------------------------
# The chaining DP (wgdi/collinearity.py:72-115)
# --------------------------------------------------------------------------------------


def _row_index(loc1: np.ndarray):
    """start offset of every distinct row value in a (loc1, loc2)-sorted point list."""
    n = len(loc1)
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(1, np.int64)
    starts = np.flatnonzero(np.r_[True, loc1[1:] != loc1[:-1]])
    return loc1[starts], np.r_[starts, n]


def chain_dp(loc1, loc2, grading, mg1, mg2, direction):
    """One direction of `score_matrix` on points sorted by (loc1, loc2) ascending.

    direction = +1 restates wgdi/collinearity.py:74-93 (sources in (loc1, loc2) order,
    window row < loc1_k < row+mg1, col < loc2_k < col+mg2, all bounds exclusive);
    direction = -1 restates wgdi/collinearity.py:96-115 (sources in (loc1 desc, loc2 asc)
    order, window row-mg1 < loc1_k < row, candidates visited nearest row first because
    `points_reverse` is sorted loc1 descending).

    Returns (score, pred, used, depth); pred = index (in the sorted point order) of the
    last source that updated the point, -1 for path roots; depth = len(path) of the
    reference's `path1`/`path2` list.
    """
    loc1 = np.asarray(loc1, np.int64)
    loc2 = np.asarray(loc2, np.int64)
    g = np.asarray(grading, np.int64)
    n = len(loc1)
    score = g.copy()                       # get_matrix :35-36
    pred = np.full(n, -1, np.int64)
    used = np.zeros(n, bool)               # usedtimes > 0 (only ever tested as < 1, :50,:61)
    depth = np.ones(n, np.int64)           # len(path), path starts as [label] (:37)
    rows, rptr = _row_index(loc1)
    nrows = len(rows)
    row_of = np.searchsorted(rows, loc1)   # dense row id of each point

    if direction > 0:
        src_order = range(n)
    else:
        # (loc1 desc, loc2 asc): walk dense rows backwards, each row left to right
        src_order = [i for r in range(nrows - 1, -1, -1) for i in range(rptr[r], rptr[r + 1])]

    for i in src_order:
        row, col = loc1[i], loc2[i]
        gap, row_old = mg2, row            # :81 / :103
        ri = row_of[i]
        stop = False
        rr = ri + direction
        while 0 <= rr < nrows and not stop:
            rk = rows[rr]
            dr = (rk - row) * direction
            if dr >= mg1:                  # exclusive window bound (:78 / :100)
                break
            lo, hi = rptr[rr], rptr[rr + 1]
            k = lo + np.searchsorted(loc2[lo:hi], col, side="right")   # loc2_k > col
            while k < hi and loc2[k] < col + mg2:                      # loc2_k < col + mg2
                dc = loc2[k] - col
                # :83-84 / :105-106 — `break`, not `continue`
                if dc > gap and (rk > row_old if direction > 0 else rk < row_old):
                    stop = True
                    break
                s = g[k] - (dr + dc)       # gap_penalty = -1 (:85 / :107)
                if s > 0 and score[k] < score[i] + s:   # strict <, first writer wins ties
                    score[k] = score[i] + s
                    pred[k] = i
                    used[i] = True
                    used[k] = True
                    depth[k] = depth[i] + 1
                    gap = min(dc, gap)     # :92 / :114
                    row_old = rk           # :93 / :115
                k += 1
            rr += direction
    return score, pred, used, depth


# --------------------------------------------------------------------------------------
# Block extraction + p-value (wgdi/collinearity.py:42-70, 117-153)
# --------------------------------------------------------------------------------------


def extract_blocks(loc1, loc2, score, pred, used, depth, order, times, params: IclParams):
    """Greedy extraction for one direction.

    `order` is nargsort_desc(score) over ALL n points (collinearity.py:49/:60); points
    with usedtimes < 1 are dropped afterwards (:50/:61).  Restates max_path (:117-144) and
    p_value_estimated (:146-153).  `times` is shared between the two directions and is
    bumped for every accepted block, also for blocks the caller drops later.

    Returns a list of (member_indices sorted by loc1 asc, rounded p, raw p, score_of_top,
    N, N1) in acceptance order, already filtered by `p > pvalue` (collinearity.py:55-56).
    """
    n = len(loc1)
    present = used.copy()
    out = []
    g0 = params.grading[0]
    for top in order:
        if not present[top]:
            continue
        # path = pred chain root..top ; rem = members still in the frame (:124-125)
        L = int(depth[top])
        rem = []
        q = top
        while q >= 0:
            if present[q]:
                rem.append(q)
            q = pred[q]
        m = len(rem)
        if L >= params.over_gap and m / L > params.coverage_ratio:      # :129
            present[rem] = False                                         # :130
            rem = np.asarray(rem, np.int64)
            l1min, l1max = loc1[rem].min(), loc1[rem].max()              # :131
            l2min, l2max = loc2[rem].min(), loc2[rem].max()
            box = (l1min <= loc1) & (loc1 <= l1max) & (l2min <= loc2) & (loc2 <= l2max)   # :134-137
            N1 = int(times[box].sum())                                   # :148
            N = int(box.sum())                                           # :149
            times[box] += 1                                              # :150
            L1 = int(l1max - l1min + 1)
            L2 = int(l2max - l2min + 1)
            sc = int(score[top])
            # :152 evaluated left to right in float64
            a = np.float64(1.0) - np.float64(sc) / np.float64(m) / np.float64(g0)
            a = a * np.float64(N1 - m + 1)
            a = a / np.float64(N)
            a = a * np.float64(L1 - m + 1)
            a = a * np.float64(L2 - m + 1)
            a = a / np.float64(L1)
            a = a / np.float64(L2)
            p = round(np.float64(a), 4)                                  # :153
            if p > params.pvalue:                                        # :55-56
                continue
            members = rem[np.argsort(loc1[rem], kind="stable")]          # :140 (loc1 unique on a path)
            out.append((members, p, float(a), sc, N, N1))
        else:
            present[top] = False                                         # :143
    return out


def collinearity_run(loc1, loc2, grading, options=(), order_fn=nargsort_desc):
    """The inner seam: `collinearity(options, points).run()` (wgdi/collinearity.py:42-70).

    Inputs are the three columns of `points` already sorted by (loc1, loc2) as
    `single_pool` does (wgdi/run_colliearity.py:164-166).  Returns
    [(loc1[], loc2[], p_value, score, direction)] — forward blocks first, then the
    anti-diagonal ones, each in acceptance order, blocks of length 3-4 included.
    """
    params = options if isinstance(options, IclParams) else IclParams(options)
    loc1 = np.asarray(loc1, np.int64)
    loc2 = np.asarray(loc2, np.int64)
    g = np.asarray(grading, np.int64)
    times = np.ones(len(loc1), np.int64)                 # get_matrix :34
    data = []
    for direction in (+1, -1):
        score, pred, used, depth = chain_dp(loc1, loc2, g, params.mg1, params.mg2, direction)
        order = order_fn(score.astype(np.int64))
        for members, p, a, sc, N, N1 in extract_blocks(loc1, loc2, score, pred, used, depth,
                                                       order, times, params):
            data.append((loc1[members], loc2[members], p, sc, direction))
    return data


# --------------------------------------------------------------------------------------
# Ingest and grading (wgdi/base.py:172-206, wgdi/run_colliearity.py:59-86, 96-106)
# --------------------------------------------------------------------------------------


def load_conf(path, section="collinearity"):
    """wgdi/base.py:27-31."""
    conf = configparser.ConfigParser()
    conf.read(path)
    return conf.items(section)


def read_lens(path):
    """Chromosome whitelist = first column as str (wgdi/base.py:198-206; only the index is used)."""
    import pandas as pd
    t = pd.read_csv(path, sep="\t", header=None, index_col=0)
    return [str(x) for x in t.index]


def read_gff(path):
    """wgdi/base.py:186-195: columns chr, id, start, end, strand, order."""
    import pandas as pd
    t = pd.read_csv(path, sep="\t", header=None)
    return {
        "id": t[1].astype(str).to_numpy(),
        "chr": t[0].astype(str).to_numpy(),
        "strand": t[4].astype(str).to_numpy(),
        "order": t[5].astype(int).to_numpy(),
    }


def grade_hits(q_codes, bitscore, multiple, repeat_number, grading):
    """`deal_blast_for_genomes` (wgdi/run_colliearity.py:59-86).

    q_codes: integer code of the query id per surviving hit (file order); bitscore: the
    column-11 values with the dtype pandas inferred.  For every query (all target
    chromosomes together) hits are ordered by nargsort_desc(bitscore) over the group's
    rows in file order; rank < multiple -> grading[0]; rank < multiple+4 -> grading[1];
    rank < repeat_number -> grading[2]; the rest are dropped (grading 0).
    """
    q_codes = np.asarray(q_codes)
    out = np.zeros(len(q_codes), np.int64)
    rednum, bluenum = int(multiple), 4 + int(multiple)
    order = np.argsort(q_codes, kind="stable")
    bounds = np.flatnonzero(np.r_[True, q_codes[order][1:] != q_codes[order][:-1], True])
    for a, b in zip(bounds[:-1], bounds[1:]):
        rows = order[a:b]                                   # group rows, file order
        ranked = rows[nargsort_desc(bitscore[rows])][:repeat_number]
        red, blue, gray = ranked[:rednum], ranked[rednum:bluenum], ranked[bluenum:repeat_number]
        out[red] = grading[0]
        out[blue] = grading[1]
        out[gray] = grading[2]
    return out


def color_hits(q_codes, bitscore, rednum, repeat_number):
    """`dotplot.pair_positon` (wgdi/dotplot.py:31-56): colour per hit, 1 red / 2 blue / 3 gray / 0 not drawn.

    :36-37  per query (groupby column 0) `sort_values(by=[11], ascending=[False])[:repeat_number]` — a one-key sort, i.e.
            nargsort_desc of the group's scores in file order;
    :35,38-40  red = first rednum, blue = [rednum : 5 + rednum], gray = [5 + rednum : repeat_number];
    :53-56  rows without a colour are dropped (`str.contains` of a word character)."""
    q_codes = np.asarray(q_codes)
    out = np.zeros(len(q_codes), np.int8)
    rednum = int(rednum)
    bluenum = 5 + rednum
    order = np.argsort(q_codes, kind="stable")
    bounds = np.flatnonzero(np.r_[True, q_codes[order][1:] != q_codes[order][:-1], True])
    for a, b in zip(bounds[:-1], bounds[1:]):
        rows = order[a:b]
        ranked = rows[nargsort_desc(bitscore[rows])][:repeat_number]
        out[ranked[:rednum]] = 1
        out[ranked[rednum:bluenum]] = 2
        out[ranked[bluenum:repeat_number]] = 3
    return out


def grade_hits_chromosomes(id1, subj, pair, bitscore, multiple, repeat_number, grading):
    """`deal_blast_for_chromosomes` (wgdi/run_colliearity.py:42-57) as pandas < 3 ran it (under pandas 3 the
    reference raises KeyError 'chr1' at :118; the golden fixture was generated with the pandas-2 behaviour of
    groupby.apply restored, tests/make_golden.py).

    id1: query id STRINGS per surviving hit (file order); subj / pair: integer codes of the subject gene and of the
    chromosome pair.  Returns (grading per hit, 0 = dropped; the row order of the frame the function returns up to
    the irrelevant concatenation order of the chromosome pairs).
      :44  stable multi-key sort by (id1 ascending, bitscore descending)
      :46  cumcount per subject gene inside each chromosome pair, in that order
      :47  keep cumcount <= repeat_number
      :48-53  pd.cut(cumcount, bins=[-1, 0, multiple, repeat_number], labels=grading, right=True)
    """
    bitscore = np.asarray(bitscore)
    n = len(bitscore)
    rn, blue = int(repeat_number), int(multiple)          # bluenum = rednum (:43)
    if not (0 < blue < rn):
        raise ValueError("bins must increase monotonically.")                  # what pd.cut raises
    if len(set(grading)) != len(grading):
        raise ValueError("labels must be unique if ordered=True; pass ordered=False for duplicate labels")
    _, id_rank = np.unique(np.asarray(id1).astype(str), return_inverse=True)
    order = np.lexsort((-bitscore, id_rank))               # stable: ties keep file order
    cum = np.zeros(n, np.int64)
    seen = {}
    for r in order.tolist():
        key = (int(pair[r]), int(subj[r]))
        c = seen.get(key, 0)
        cum[r] = c
        seen[key] = c + 1
    grade = np.zeros(n, np.int64)
    grade[cum == 0] = grading[0]
    grade[(cum > 0) & (cum <= blue)] = grading[1]
    grade[(cum > blue) & (cum <= rn)] = grading[2]
    return grade, order


def icl_text(conf_path):
    """The outer seam: everything `mycollinearity(options).run()` writes to `savefile`
    (wgdi/run_colliearity.py:88-200), returned as a string instead of written."""
    import pandas as pd
    opts = dict(load_conf(conf_path))
    options = list(opts.items())
    grading = [int(k) for k in opts["grading"].split(",")] if "grading" in opts else [50, 40, 25]
    repeat_number = int(opts.get("repeat_number", 10))
    multiple = int(opts.get("multiple", 1))
    min_score = int(opts.get("score", 100))
    max_evalue = float(opts.get("evalue", 1e-5))
    over_gap = 5                                             # run_colliearity.py:21
    params = IclParams(options)

    chroms1, chroms2 = read_lens(opts["lens1"]), read_lens(opts["lens2"])
    gff1, gff2 = read_gff(opts["gff1"]), read_gff(opts["gff2"])

    def keep(gff, chroms):                                   # run_colliearity.py:96-97
        m = np.isin(gff["chr"], np.asarray(chroms, dtype=object).astype(str))
        return {k: v[m] for k, v in gff.items()}
    gff1, gff2 = keep(gff1, chroms1), keep(gff2, chroms2)
    idx1 = {g: i for i, g in enumerate(gff1["id"])}
    idx2 = {g: i for i, g in enumerate(gff2["id"])}

    blast = pd.read_csv(opts["blast"], sep="\t", header=None)            # base.py:174
    id1 = blast[0].astype(str).to_numpy()
    id2 = blast[1].astype(str).to_numpy()
    if opts.get("blast_reverse", False) == "true":                      # base.py:176-177
        id1, id2 = id2, id1
    bits = blast[11].to_numpy()
    ev = blast[10].to_numpy()
    m = (bits >= min_score) & (ev < max_evalue) & (id1 != id2)          # base.py:178
    g1 = np.array([idx1.get(x, -1) for x in id1])
    g2 = np.array([idx2.get(x, -1) for x in id2])
    m &= (g1 >= 0) & (g2 >= 0)                                          # base.py:179
    rows = np.flatnonzero(m)
    seen, keep_rows = set(), []
    for r in rows:                                                      # base.py:180 keep='first'
        key = (g1[r], g2[r])
        if key not in seen:
            seen.add(key)
            keep_rows.append(r)
    rows = np.asarray(keep_rows, np.int64)
    g1, g2, bits = g1[rows], g2[rows], bits[rows]
    # groupby([0]) iterates query ids in sorted order but gradings are set by label, so the
    # iteration order is irrelevant; only the per-group file order matters.
    comparison = str(opts.get("comparison", "genomes")).lower()         # :22, :108-111
    if comparison == "chromosomes":
        c1u, c1c = np.unique(gff1["chr"][g1], return_inverse=True)
        c2u, c2c = np.unique(gff2["chr"][g2], return_inverse=True)
        grade, order = grade_hits_chromosomes(id1[rows], g2, c1c * len(c2u) + c2c, bits, multiple, repeat_number, grading)
        g1, g2, grade = g1[order], g2[order], grade[order]             # the rows now follow the sort of :44
    elif comparison == "genomes":
        grade = grade_hits(g1, bits, multiple, repeat_number, grading)
    else:
        return None                                                     # neither branch runs; the reference then fails at :118
    ok = grade > 0                                                      # run_colliearity.py:86
    g1, g2, grade = g1[ok], g2[ok], grade[ok]
    if len(g1) < 1:
        return None
    loc1, loc2 = gff1["order"][g1], gff2["order"][g2]                   # :103-104
    chr1, chr2 = gff1["chr"][g1], gff2["chr"][g2]                       # :105-106

    text = []
    num = 1
    pairs = sorted(set(zip(chr1.tolist(), chr2.tolist())))              # groupby sorted string keys, :118
    name1 = {}
    for c in set(chr1.tolist()):
        mm = gff1["chr"] == c
        name1[c] = (dict(zip(gff1["order"][mm], gff1["id"][mm])), dict(zip(gff1["order"][mm], gff1["strand"][mm])))
    name2 = {}
    for c in set(chr2.tolist()):
        mm = gff2["chr"] == c
        name2[c] = (dict(zip(gff2["order"][mm], gff2["id"][mm])), dict(zip(gff2["order"][mm], gff2["strand"][mm])))
    for c1, c2 in pairs:
        sel = np.flatnonzero((chr1 == c1) & (chr2 == c2))
        o = sel[np.lexsort((loc2[sel], loc1[sel]))]                     # :164-166 stable (loc1, loc2)
        data = collinearity_run(loc1[o], loc2[o], grade[o], params)
        for b1, b2, p, sc, _d in data:
            if len(b1) < over_gap:                                      # :178
                continue
            mark = "plus" if b2[0] < b2[1] else "minus"                 # :194-195
            # header is re-cut at the first ':' pair by the parent (:147-151)
            tail = f" score={sc} pvalue={p} N={len(b1)} {c1}&{c2} {mark}"
            text.append(f"# Alignment {num}:" + tail.split(":")[0])
            num += 1
            n1, s1 = name1[c1]
            n2, s2 = name2[c2]
            for x, y in zip(b1.tolist(), b2.tolist()):
                strand = "1" if s1[x] == s2[y] else "-1"               # :185-187
                text.append(f"{n1[x]} {x} {n2[y]} {y} {strand}")
    return "\n".join(text) + "\n" if text else ""
