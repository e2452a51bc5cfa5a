"""Drop-in for the reference's `-bi` tool: `block_info(options).run()` (wgdi/block_info.py:6-165) for collinearity
files in WGDI's own format (what `-icl` writes).  Same `[blockinfo]` options, same printed lines, same CSV.

Where the reference loops over blocks and gene pairs in Python — one pandas lookup per gene pair and two `isin`
scans of the whole BLAST table per block — the per-block integer statistics come from one CUDA launch over all
blocks (include/icl_blockinfo.h); file parsing, the per-query hit rank (numpy's own sort where ties matter) and
the CSV are host work.  Extra optional key: `device = <cuda index>`.
"""
from __future__ import annotations

import os
import re
import time

import numpy as np
import pandas as pd

from . import _bi, _host, base
from .ingest import rank_hits, rank_hits_native


def str_to_bool(s):
    """wgdi/base.py:276-279."""
    if isinstance(s, bool):
        return s
    return str(s).strip().lower() == 'true'


def get_median(data):
    """wgdi/base.py:154-160."""
    if not data:
        return 0
    data_sorted = sorted(data)
    half = len(data_sorted) // 2
    return (data_sorted[half] + data_sorted[-(half + 1)]) / 2


def read_ks(file, col):
    """wgdi/base.py:144-151 — Series 'id1,id2' -> Ks (first occurrence of a pair, values >= 0)."""
    ks = pd.read_csv(file, sep='\t')
    ks = ks.drop_duplicates(subset=['id1', 'id2'], keep='first')
    ks[col] = ks[col].astype(float)
    ks = ks[ks[col] >= 0]
    ks.index = ks['id1'] + ',' + ks['id2']
    return ks[col]


_ODD_SPACE = re.compile(r"[^\S \n]")      # whitespace other than ' ' and the line ends
_HEADER_NUM = re.compile(r'[\.\d+]+')


def read_collinearity_flat(fn):
    r"""wgdi/base.py:120-141 with the blocks flattened: (ids, pvalues, offsets, name1, loc1, name2, loc2) where
    ids / pvalues are the header strings `arr[0]` / `arr[2]` and block b owns members offsets[b]..offsets[b+1].
    The reference splits member lines with re.split(r"\s", line): one split per whitespace character.  When the
    file holds no whitespace besides ' ' and newlines that is str.split(' '), which is what the fast path uses."""
    with open(fn) as f:
        text = f.read()
    plain = _ODD_SPACE.search(text) is None
    ids, pvs, off = [], [], [0]
    n1, l1, n2, l2 = [], [], [], []
    flag, arr, count = 0, [], 0
    lines = text.split('\n')
    if lines and lines[-1] == '':                                  # file iteration yields no line after the last newline
        lines.pop()
    for line in lines:
        line = line.strip()
        if line.startswith('# Alignment'):                         # re.match(r"# Alignment", line)
            flag = 1
            if count:
                ids.append(arr[0]); pvs.append(arr[2]); off.append(off[-1] + count)
                count = 0
            arr = _HEADER_NUM.findall(line)
            continue
        if flag == 0:
            continue
        k = line.split(' ') if plain else re.split(r"\s", line)
        n1.append(k[0]); l1.append(k[1]); n2.append(k[2]); l2.append(k[3])
        count += 1
    if count:
        ids.append(arr[0]); pvs.append(arr[2]); off.append(off[-1] + count)
    return ids, pvs, np.asarray(off, np.int64), n1, l1, n2, l2


def frame_to_csv(df, path):
    """`df.to_csv(path, index=None)` (wgdi/block_info.py:75) for the column types this frame has, several times faster:
    pandas turns every float column into text with numpy's `astype(str)` and sends every row through the csv module; here
    floats are printed with `repr` (the same shortest round-trip text), ints with `str`, missing values as '', rows
    are joined directly.  Anything the shortcut does not cover — a cell that needs quoting, another column type, a
    separator inside a column name — is written by pandas itself.  Tested against `DataFrame.to_csv` byte for byte."""
    special = (',', '"', '\n', '\r')
    cols = []
    ok = not any(ch in str(name) for name in df.columns for ch in special)
    for name in df.columns if ok else ():
        c = df[name]
        kind = getattr(c.dtype, "kind", "O")
        if c.dtype == np.float64:
            v = c.to_numpy()
            out = list(map(repr, v.tolist()))
            for i in np.flatnonzero(np.isnan(v)).tolist():
                out[i] = ''
        elif kind in "iu" and isinstance(c.dtype, np.dtype):
            out = list(map(str, c.to_numpy().tolist()))
        elif kind == "O" or str(c.dtype) in ("str", "string", "object"):
            vals = c.tolist()
            if not {type(v) for v in vals} <= {str, float, int, type(None)}:
                ok = False
                break
            out = ['' if (v is None or v != v) else (v if type(v) is str else repr(v) if type(v) is float else str(v))
                   for v in vals]
            joined = "\x00".join(out)
            if any(ch in joined for ch in special):
                ok = False
                break
        else:
            ok = False
            break
        cols.append(out)
    if not ok:
        df.to_csv(path, index=None)
        return
    lines = [",".join(map(str, df.columns))]
    lines.extend(map(",".join, zip(*cols)))
    with open(path, "w", newline="") as f:
        f.write("\n".join(lines) + "\n")


class block_info():
    def __init__(self, options):
        self.repeat_number = 20
        self.ks_col = 'ks_NG86'
        self.blast_reverse = False
        for k, v in options:
            setattr(self, str(k), v)
            print(f"{k} = {v}")
        self.repeat_number = int(self.repeat_number)
        self.blast_reverse = str_to_bool(self.blast_reverse)
        self.device = int(getattr(self, 'device', 0))
        self.native = getattr(self, 'ingest', 'native').lower() != 'pandas' and _host.available()
        self.timing = {}

    def auto_file(self, ids1=None, ids2=None):
        """wgdi/block_info.py:167-180: only WGDI's own format is offered here.  With the packed gene tables (native host
        library, no Ks file) the file is read by icl_host_read_collinearity — a dict, see `_host.read_collinearity` —
        unless it is not the plain text -icl writes; then, and otherwise, `read_collinearity_flat`'s tuple."""
        with open(self.collinearity) as f:
            p = ' '.join(f.readlines()[0:30])
        if 'path length' in p or 'MAXIMUM GAP' in p or 'MATCH_SIZE' in p or '## Alignment' in p:
            raise NotImplementedError("only WGDI-format collinearity files (as written by -icl) are supported")
        if '# Alignment' in p:
            if ids1 is not None:
                nat = _host.read_collinearity(self.collinearity, ids1, ids2)
                if nat is not None:
                    return nat
            return read_collinearity_flat(self.collinearity)
        raise NotImplementedError("only WGDI-format collinearity files (as written by -icl) are supported")

    def load_hits(self, gff1, gff2, ids1=None, ids2=None):
        """base.newblast + blast_homo's trimming (block_info.py:149-150, :104-105): kept hits with their rank.
        The id columns are NEVER swapped here, whatever `blast_reverse` says: the reference turns the option into a bool
        (wgdi/block_info.py:16) and hands that bool to base.newblast, whose test is `reverse == 'true'` (base.py:176) —
        False for either bool.  Reproduced as is (drop-in; pinned by test_block_info_ignores_blast_reverse)."""
        if self.native and not _host.is_compressed(self.blast):
            if ids1 is None:
                ids1 = _host.pack_strings(gff1.index)
                ids2 = ids1 if self.gff2 == self.gff1 and self.lens2 == self.lens1 else _host.pack_strings(gff2.index)
            hits = _host.read_m8(self.blast, False, int(self.score), float(self.evalue), ids1, ids2)
            g1, g2, bits = hits['g1'], hits['g2'], hits['bits']
        else:
            blast = base.newblast(self.blast, int(self.score), float(self.evalue), gff1, gff2, self.blast_reverse)
            g1, g2, bits = blast['g1'].to_numpy(), blast['g2'].to_numpy(), blast[11].to_numpy()
        rank = rank_hits_native(g1, bits, len(gff1), self.repeat_number) if self.native else rank_hits(g1, bits, self.repeat_number)
        keep = rank >= 0
        return g1[keep], g2[keep], rank[keep]

    def run(self):
        t0 = time.perf_counter()
        lens1 = base.newlens(self.lens1, self.position)
        lens2 = base.newlens(self.lens2, self.position)
        same = self.gff2 == self.gff1 and self.lens2 == self.lens1     # self comparison: the tables are parsed once
        gff1 = base.newgff(self.gff1, usecols=range(6))
        gff1 = gff1[gff1['chr'].isin(lens1.index)]
        gff2 = gff1 if same else base.newgff(self.gff2, usecols=range(6))
        if not same:
            gff2 = gff2[gff2['chr'].isin(lens2.index)]
        ids1 = ids2 = None
        if self.native:
            ids1 = _host.pack_strings(gff1.index)
            ids2 = ids1 if same else _host.pack_strings(gff2.index)
        g1, g2, rank = self.load_hits(gff1, gff2, ids1, ids2)
        order2 = gff2['order'].to_numpy()
        o = np.lexsort((rank, g1))
        hit_off = np.zeros(len(gff1) + 1, np.int64)
        np.cumsum(np.bincount(g1, minlength=len(gff1)), out=hit_off[1:])
        ks = pd.Series([], dtype=float) if (not hasattr(self, 'ks') or self.ks == 'none' or self.ks == '') \
            else read_ks(self.ks, self.ks_col)
        # the Ks columns are looked up by gene NAME (:77-86): with a Ks table the Python reader keeps the names
        parsed = self.auto_file(ids1, ids2) if len(ks) == 0 else self.auto_file()
        native = parsed if isinstance(parsed, dict) else None
        if native is not None:
            ids, pvs, off, m1, m2 = native['ids'], native['pvalues'], native['off'], native['g1'], native['g2']
            n1 = l1 = n2 = l2 = None
        else:
            ids, pvs, off, n1, l1, n2, l2 = parsed
            m1 = gff1.index.get_indexer(pd.Index(n1, dtype=object)) if n1 else np.zeros(0, np.int64)
            m2 = gff2.index.get_indexer(pd.Index(n2, dtype=object)) if n2 else np.zeros(0, np.int64)
        mo2 = np.where(m2 >= 0, order2[np.maximum(m2, 0)], 0) if len(m2) else np.zeros(0, np.int64)
        t1 = time.perf_counter()
        homo_sum, homo_cnt, t_rows, t_uniq = _bi.block_stats(self.device, hit_off, g2[o], rank[o], order2[g2[o]] if len(o) else [],
                                                             off, m1, m2, mo2, self.repeat_number)
        t2 = time.perf_counter()
        data = self.assemble(ids, pvs, off, n1, l1, n2, l2, m1, m2, gff1, gff2, ks, homo_sum, homo_cnt, t_rows, t_uniq, native)
        data['class1'] = 0
        data['class2'] = 0
        frame_to_csv(data, self.savefile)                       # data.to_csv(self.savefile, index=None), :75
        t3 = time.perf_counter()
        self.timing.update(ingest_s=t1 - t0, device_s=t2 - t1, write_s=t3 - t2, blocks=len(ids))

    def assemble(self, ids, pvs, off, n1, l1, n2, l2, m1, m2, gff1, gff2, ks, homo_sum, homo_cnt, t_rows, t_uniq, native=None):
        """The frame of block_position (block_info.py:18-75), built from the same Python values in the same order so
        that pandas infers the same column types and `to_csv` prints the same text."""
        chr1_of, chr2_of = gff1['chr'].to_numpy(), gff2['chr'].to_numpy()
        have_ks = len(ks) > 0
        ksd = dict(zip(ks.index.tolist(), ks.to_numpy())) if have_ks else {}
        # float(k[1]) / int(float(k[1])) / str(...) of :31, :37-38, :51-52 for all members at once
        if native is not None:                                       # the same values, formed by icl_host_read_collinearity
            f1, f2, blk1, blk2 = native['loc1'].tolist(), native['loc2'].tolist(), native['block1'], native['block2']
        else:
            f1 = np.asarray(l1, dtype=np.float64) if l1 else np.zeros(0)
            f2 = np.asarray(l2, dtype=np.float64) if l2 else np.zeros(0)
            s1 = f1.astype(np.int64).astype(str).tolist()
            s2 = f2.astype(np.int64).astype(str).tolist()
            f1, f2 = f1.tolist(), f2.tolist()
            blk1 = blk2 = None
        # :98-100 mean over the pairs that are kept hits; :134 tandem ratio — same IEEE operations as the Python floats
        Ls = np.diff(off).astype(np.int64)
        with np.errstate(divide='ignore', invalid='ignore'):
            hmean = (homo_sum.astype(np.float64) / homo_cnt[:, None]).tolist()
            tandem = (t_uniq / Ls * t_rows / (Ls + t_rows)).tolist() if len(Ls) else []
        if native is not None and not have_ks and os.environ.get("ICL_BI_ROWS", "0") != "1":
            frame = self._assemble_columns(ids, pvs, off, m1, m2, chr1_of, chr2_of, native, homo_sum, homo_cnt, tandem)
            if frame is not None:
                return self._with_densities(frame)
        cnts = homo_cnt.tolist()
        none5 = [-1, -1, -1, -1, -1]
        data = []
        for b in range(len(ids)):
            a0, a1 = int(off[b]), int(off[b + 1])
            if m1[a0] < 0 or m2[a0] < 0:                             # :24-25
                continue
            L = a1 - a0
            if have_ks:                                              # :40-41, :77-86
                blk_ks = []
                for x, y in zip(n1[a0:a1], n2[a0:a1]):
                    v = ksd.get(f"{x},{y}")
                    if v is None:
                        v = ksd.get(f"{y},{x}", -1)
                    blk_ks.append(v)
                ks_arr = [k for k in blk_ks if k >= 0]               # :88-96
                ks_median, ks_average = (get_median(ks_arr), sum(ks_arr) / len(ks_arr)) if ks_arr else (-1, -1)
                blkks = '_'.join([str(k) for k in blk_ks])
            else:
                ks_median, ks_average, blkks = -1, -1, '_'.join(['-1'] * L)
            homo = hmean[b] if cnts[b] > 0 else none5
            data.append([ids[b], chr1_of[m1[a0]], chr2_of[m2[a0]], f1[a0], f1[a1 - 1], f2[a0], f2[a1 - 1], pvs[b], L,
                         ks_median, ks_average, *homo, blk1[b] if blk1 is not None else '_'.join(s1[a0:a1]),
                         blk2[b] if blk2 is not None else '_'.join(s2[a0:a1]), blkks, tandem[b]])
        data_df = pd.DataFrame(data, columns=[
            'id', 'chr1', 'chr2', 'start1', 'end1', 'start2', 'end2', 'pvalue', 'length',
            'ks_median', 'ks_average', 'homo1', 'homo2', 'homo3', 'homo4', 'homo5',
            'block1', 'block2', 'ks', 'tandem_ratio'])
        return self._with_densities(data_df)

    @staticmethod
    def _with_densities(data_df):
        data_df['density1'] = data_df['length'] / ((data_df['end1'] - data_df['start1']).abs() + 1)     # :71-72
        data_df['density2'] = data_df['length'] / ((data_df['end2'] - data_df['start2']).abs() + 1)
        return data_df

    @staticmethod
    def _assemble_columns(ids, pvs, off, m1, m2, chr1_of, chr2_of, native, homo_sum, homo_cnt, tandem):
        """`assemble` without its per-block loop, for the common case (native reader, no Ks table): the same values column
        by column.  pandas infers a column's type from its values, so a column that the row loop would fill with the int -1
        only stays int64 and one that mixes -1 with means becomes float64 — as there.  None when no block survives (the row
        loop handles that frame)."""
        a0, a1 = off[:-1], off[1:]
        vb = np.flatnonzero((m1[a0] >= 0) & (m2[a0] >= 0)) if len(a0) else np.zeros(0, np.int64)    # :24-25
        if len(vb) == 0:
            return None
        a0, a1 = a0[vb], a1[vb]
        f1, f2 = native['loc1'], native['loc2']
        L = (a1 - a0).astype(np.int64)
        cnt = homo_cnt[vb]
        with np.errstate(divide='ignore', invalid='ignore'):
            hmean = homo_sum[vb].astype(np.float64) / cnt[:, None]
        some = bool((cnt > 0).any())
        vbl = vb.tolist()
        ks_text = {n: '_'.join(['-1'] * n) for n in set(L.tolist())}
        cols = {
            'id': [ids[b] for b in vbl], 'chr1': chr1_of[m1[a0]].tolist(), 'chr2': chr2_of[m2[a0]].tolist(),
            'start1': f1[a0], 'end1': f1[a1 - 1], 'start2': f2[a0], 'end2': f2[a1 - 1],
            'pvalue': [pvs[b] for b in vbl], 'length': L,
            'ks_median': np.full(len(vb), -1, np.int64), 'ks_average': np.full(len(vb), -1, np.int64)}
        for j in range(5):
            cols[f'homo{j + 1}'] = np.where(cnt > 0, hmean[:, j], -1.0) if some else np.full(len(vb), -1, np.int64)
        cols.update({'block1': [native['block1'][b] for b in vbl], 'block2': [native['block2'][b] for b in vbl],
                     'ks': [ks_text[n] for n in L.tolist()], 'tandem_ratio': np.asarray(tandem, np.float64)[vb]})
        return pd.DataFrame(cols)
