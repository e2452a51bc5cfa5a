"""CPU restatement of `wgdi -bi` (block information; wgdi/block_info.py) for WGDI-format collinearity files.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): the product path (wgdi_b200/block_info.py + the CUDA
kernel of wgdi_b200/csrc/icl_blockinfo.cu) never imports it.  Plain Python loops, line-by-line citations;
pinned against files the unmodified reference wrote (tests/make_golden.py -> tests/golden/e2e_*/expected.blockinfo*.csv).

What the reference computes per block of the collinearity file (block_info.py:18-75):
  chr1, chr2, start/end on both genomes, the header's p value, the length, Ks median/average of its gene pairs,
  homo1..homo5 = mean over the block's gene pairs of a {1, 0, -1} class derived from the pair's per-query BLAST
  rank (blast_homo :102-120), the gene-order lists, and a tandem ratio (:122-134).
"""
from __future__ import annotations

import io
import re

import numpy as np

from .icl_oracle import load_conf, nargsort_desc, read_gff, read_lens  # noqa: F401  (same readers as -icl)


def read_collinearity(path):
    """wgdi/base.py:120-141: [[id, [[name1, loc1, name2, loc2, strand], ...], pvalue], ...]; the header numbers are
    the runs of `[.\\d+]` characters, [0] = alignment id, [2] = p value (strings)."""
    data, b, flag, arr = [], [], 0, []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if re.match(r"# Alignment", line):
                flag = 1
                if not b:
                    arr = re.findall(r"[\.\d+]+", line)
                    continue
                data.append([arr[0], b, arr[2]])
                b = []
                arr = re.findall(r"[\.\d+]+", line)
                continue
            if flag == 0:
                continue
            b.append(re.split(r"\s", line))
        if b:
            data.append([arr[0], b, arr[2]])
    return data


def read_ks(path, col):
    """wgdi/base.py:144-151: first occurrence of every (id1, id2), values >= 0, keyed 'id1,id2'."""
    import pandas as pd
    ks = pd.read_csv(path, sep="\t")
    ks = ks.drop_duplicates(subset=["id1", "id2"], keep="first")
    v = ks[col].astype(float)
    ok = (v >= 0).to_numpy()
    keys = (ks["id1"].astype(str) + "," + ks["id2"].astype(str)).to_numpy()[ok]
    return dict(zip(keys.tolist(), v.to_numpy()[ok]))            # values stay numpy.float64 like Series items


def get_median(data):
    """wgdi/base.py:154-160."""
    if not data:
        return 0
    s = sorted(data)
    half = len(s) // 2
    return (s[half] + s[-(half + 1)]) / 2


def filtered_blast(opts, gff1, gff2):
    """base.newblast (wgdi/base.py:172-183): surviving rows in file order -> (id1, id2, bitscore as parsed)."""
    import pandas as pd
    blast = pd.read_csv(opts["blast"], sep="\t", header=None)
    id1 = blast[0].astype(str).to_numpy()
    id2 = blast[1].astype(str).to_numpy()
    if str(opts.get("blast_reverse", "false")).strip().lower() == "true":   # block_info.py:15 str_to_bool, then base.py:176
        id1, id2 = id2, id1
    bits, ev = blast[11].to_numpy(), blast[10].to_numpy()
    m = (bits >= int(opts["score"])) & (ev < float(opts["evalue"])) & (id1 != id2)
    s1, s2 = set(gff1["id"].tolist()), set(gff2["id"].tolist())
    seen, keep = set(), []
    for r in np.flatnonzero(m).tolist():
        if id1[r] in s1 and id2[r] in s2 and (id1[r], id2[r]) not in seen:
            seen.add((id1[r], id2[r]))
            keep.append(r)
    keep = np.asarray(keep, np.int64)
    return id1[keep], id2[keep], bits[keep]


def blast_homo(id1, id2, bits, repeat_number):
    """block_info.py:102-120: per query, hits ordered by pandas' sort_values(by=11, ascending=False) (numpy's unstable
    argsort, hazard H2), the first `repeat_number` kept.  homo_i = 1 for rank < i, 0 for i <= rank < i + 5,
    -1 for i + 5 <= rank < repeat_number.  Returns {(id1, id2): rank} for the kept hits and, per query, its kept
    (id2, rank) list."""
    rank = {}
    groups = {}
    for r, q in enumerate(id1.tolist()):
        groups.setdefault(q, []).append(r)
    for q, rows in groups.items():
        rows = np.asarray(rows)
        ranked = rows[nargsort_desc(bits[rows])][:repeat_number]
        for k, r in enumerate(ranked.tolist()):
            rank[(q, id2[r])] = k
    return rank


def homo_values(rank, repeat_number):
    out = []
    for i in range(1, 6):
        out.append(1 if rank < i else 0 if rank < i + 5 else -1)
    return out


def block_info_csv(conf_path, section="blockinfo"):
    """Everything `block_info(options).run()` writes to `savefile` (block_info.py:136-165), returned as text."""
    import pandas as pd
    opts = dict(load_conf(conf_path, section))
    repeat_number = int(opts.get("repeat_number", 20))
    ks_col = opts.get("ks_col", "ks_NG86")
    chroms1, chroms2 = read_lens(opts["lens1"]), read_lens(opts["lens2"])
    gff1, gff2 = read_gff(opts["gff1"]), read_gff(opts["gff2"])

    def keep(gff, chroms):                                                   # :145-146
        m = np.isin(gff["chr"], np.asarray(chroms, dtype=object).astype(str))
        return {k: v[m] for k, v in gff.items()}
    gff1, gff2 = keep(gff1, chroms1), keep(gff2, chroms2)
    chr_of1 = dict(zip(gff1["id"].tolist(), gff1["chr"].tolist()))
    chr_of2 = dict(zip(gff2["id"].tolist(), gff2["chr"].tolist()))
    order_of2 = dict(zip(gff2["id"].tolist(), gff2["order"].tolist()))

    id1, id2, bits = filtered_blast(opts, gff1, gff2)                        # :149
    rank = blast_homo(id1, id2, bits, repeat_number)                         # :150
    hits_of = {}                                                             # trimmed blast, per query
    for (q, t), k in rank.items():
        hits_of.setdefault(q, []).append((t, order_of2[t]))
    ks_opt = opts.get("ks", "none")
    ks = {} if ks_opt in ("none", "") else read_ks(ks_opt, ks_col)           # :157

    data = []
    for bid, rows, pvalue in read_collinearity(opts["collinearity"]):        # :154, :20
        if rows[0][0] not in chr_of1 or rows[0][2] not in chr_of2:           # :24-25
            continue
        chr1, chr2 = chr_of1[rows[0][0]], chr_of2[rows[0][2]]
        a1, a2 = [float(k[1]) for k in rows], [float(k[3]) for k in rows]
        blk_homo, blk_ks, b1, b2 = [], [], [], []
        for k in rows:
            b1.append(int(float(k[1])))
            b2.append(int(float(k[3])))
            pair, rev = f"{k[0]},{k[2]}", f"{k[2]},{k[0]}"                   # :78-86
            blk_ks.append(ks[pair] if pair in ks else ks[rev] if rev in ks else -1)
            if (k[0], k[2]) in rank:                                         # :44-45
                blk_homo.append(homo_values(rank[(k[0], k[2])], repeat_number))
        ks_arr = [v for v in blk_ks if v >= 0]                               # :88-96
        ks_median, ks_average = (get_median(ks_arr), sum(ks_arr) / len(ks_arr)) if ks_arr else (-1, -1)
        homo = pd.DataFrame(blk_homo).mean().values if blk_homo else [-1, -1, -1, -1, -1]      # :98-100
        # tandem ratio, :122-134
        ids1 = set(k[0] for k in rows)
        ids2 = set(k[2] for k in rows)
        n_rows, uniq = 0, set()
        for k in rows:
            o2 = order_of2.get(k[2])
            if o2 is None:
                continue                                                      # NaN order2: both comparisons are False
            for t, ot in hits_of.get(k[0], ()):
                if t in ids2 and k[0] in ids1:
                    d = abs(ot - o2)
                    if 0 < d <= repeat_number:
                        n_rows += 1
                        uniq.add(k[0])
        tandem = len(uniq) / len(rows) * n_rows / (len(rows) + n_rows)
        data.append([bid, chr1, chr2, a1[0], a1[-1], a2[0], a2[-1], pvalue, len(rows), ks_median, ks_average, *homo,
                     "_".join(map(str, b1)), "_".join(map(str, b2)), "_".join(str(v) for v in blk_ks), tandem])
    df = pd.DataFrame(data, columns=["id", "chr1", "chr2", "start1", "end1", "start2", "end2", "pvalue", "length",
                                     "ks_median", "ks_average", "homo1", "homo2", "homo3", "homo4", "homo5",
                                     "block1", "block2", "ks", "tandem_ratio"])
    df["density1"] = df["length"] / ((df["end1"] - df["start1"]).abs() + 1)   # :71-72
    df["density2"] = df["length"] / ((df["end2"] - df["start2"]).abs() + 1)
    df["class1"] = 0                                                          # :161-162
    df["class2"] = 0
    buf = io.StringIO()
    df.to_csv(buf, index=None)                                                # :165
    return buf.getvalue()
