"""Host-side readers of the `-icl` path: same names and argument meaning as the helpers the
reference keeps in wgdi/base.py (load_conf :27-31, newblast :172-183, newgff :186-195,
newlens :198-206), so `wgdi_b200.run_colliearity` reads like the reference's driver.
"""
from __future__ import annotations

import configparser

import numpy as np


class _LazyPandas:
    """pandas is imported at the first read, not with this module: its import takes about a second, and the drop-in driver
    has by then started the CUDA context on a background thread (run_colliearity.mycollinearity.run) — the two seconds of a
    fresh process overlap instead of adding up."""

    def __getattr__(self, name):
        import pandas
        return getattr(pandas, name)


pd = _LazyPandas()


def load_conf(file, section):
    """wgdi/base.py:27-31 — `[(key, value_string)]` of one INI section (keys lower-cased)."""
    conf = configparser.ConfigParser()
    conf.read(file)
    return conf.items(section)


def newlens(file, position):
    """wgdi/base.py:198-206 — Series indexed by str(chromosome); column 2 for 'order', 1 for 'end'.
    The -icl path only uses the index (wgdi/run_colliearity.py:96-97)."""
    lens = pd.read_csv(file, sep="\t", header=None, index_col=0)
    lens.index = lens.index.astype(str)
    if position == "order":
        lens = lens[2]
    elif position == "end":
        lens = lens[1]
    return lens


def newgff(file, usecols=None):
    """wgdi/base.py:186-195 — frame indexed by gene id with chr (str), start, end, strand (str), order.
    usecols: the columns to parse (the -icl driver passes the six the path reads; a seventh, free-text column costs a
    third of the parse); None reads them all like the reference."""
    gff = pd.read_csv(file, sep="\t", header=None, index_col=1, usecols=usecols)
    gff.rename(columns={0: "chr", 2: "start", 3: "end", 4: "strand", 5: "order"}, inplace=True)
    gff["chr"] = gff["chr"].astype(str)
    gff["start"] = gff["start"].astype(np.int64)
    gff["end"] = gff["end"].astype(np.int64)
    gff["strand"] = gff["strand"].astype(str)
    gff["order"] = gff["order"].astype(int)
    return gff


def newblast(file, score, evalue, gene_loc1, gene_loc2, reverse):
    """wgdi/base.py:172-183 — read the m8 table and keep the rows the reference keeps:
    bitscore >= score, evalue < evalue, id1 != id2 (:178), both ids known (:179), first
    occurrence of every (id1, id2) (:180); the two id columns are swapped only when
    `reverse == 'true'` (:176-177).  Returns a frame with the original row labels, the id
    columns 0/1, column 11 (bit score, dtype as parsed) and the integer gene codes `g1`/`g2`
    (row numbers in gene_loc1 / gene_loc2), which is what the device path consumes.
    """
    blast = pd.read_csv(file, sep="\t", header=None)
    c0, c1 = (1, 0) if reverse == "true" else (0, 1)
    id1, id2 = blast[c0], blast[c1]
    keep = (blast[11] >= score) & (blast[10] < evalue) & (id2 != id1)
    g1 = gene_loc1.index.get_indexer(id1)
    g2 = gene_loc2.index.get_indexer(id2)
    keep &= (g1 >= 0) & (g2 >= 0)
    rows = np.flatnonzero(keep.to_numpy())
    g1, g2 = g1[rows], g2[rows]
    first = ~pd.Series(g1.astype(np.int64) * max(len(gene_loc2), 1) + g2).duplicated(keep="first").to_numpy()
    rows, g1, g2 = rows[first], g1[first], g2[first]
    out = pd.DataFrame({0: id1.to_numpy()[rows].astype(str), 1: id2.to_numpy()[rows].astype(str),
                        11: blast[11].to_numpy()[rows], "g1": g1, "g2": g2}, index=blast.index[rows])
    return out
