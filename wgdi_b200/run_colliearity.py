"""Drop-in for the reference's `-icl` driver: `mycollinearity(options).run()`
(wgdi/run_colliearity.py:13-200).  Same `[collinearity]` options, same printed lines, same
output file; the per-chromosome-pair Python/pandas work of the reference's worker processes
(`single_pool`, :158-200 -> wgdi/collinearity.py) runs on the GPU for all pairs at once.

Differences a user can observe: `process` is accepted and ignored (there is no worker pool:
chromosome pairs are work items of one CUDA launch, or shards over several GPUs, see
wgdi_b200/shard.py), and `comparison = chromosomes` works (the reference itself raises
KeyError there under pandas >= 3, SURVEY.md hazard H19; the semantics are those of its code
under pandas < 3, pinned by tests/golden/e2e_cross_chrom).
"""
from __future__ import annotations

import sys
import time

import numpy as np

from . import _host, base
from .engine import ChainParams, IclEngine
from .params import compact_points
from .ingest import (classes_of_ambiguous, grade_hits, grade_hits_chromosomes, grade_hits_native, pair_ids,
                     pair_table)


class mycollinearity():
    def __init__(self, options):
        # defaults of wgdi/run_colliearity.py:16-22
        self.repeat_number = 10
        self.multiple = 1
        self.score = 100
        self.evalue = 1e-5
        self.blast_reverse = False
        self.over_gap = 5
        self.comparison = 'genomes'
        self.options = options
        for k, v in options:
            setattr(self, str(k), v)
            print(f"{str(k)} = {v}")                      # :27
        self.position = 'order'                           # :28 — `position` is read and overridden
        if hasattr(self, 'grading'):
            self.grading = [int(k) for k in self.grading.split(',')]
        else:
            self.grading = [50, 40, 25]
        self.process = int(self.process) if hasattr(self, 'process') else 4
        self.over_gap = int(self.over_gap)
        self.device = int(getattr(self, 'device', 0))     # extra, optional key: CUDA device index
        # extra, optional key: `devices = 0,1,2,3` shards the chromosome pairs over several GPUs (LPT by point count)
        self.devices = [int(k) for k in str(getattr(self, 'devices', self.device)).split(',') if k.strip() != '']
        # extra, optional key: `ingest = pandas` forces the reference-like pandas readers of wgdi_b200.base;
        # the default uses the native parser / writer (libicl_host.so) when it is built.  Same results.
        # (`ingest = native-steps`: the native parser and grading as separate calls with numpy in between — the path the
        # fused ingest is tested against.)
        ingest = getattr(self, 'ingest', 'native').lower()
        self.native = ingest != 'pandas' and _host.available()
        self.fused = ingest != 'native-steps'
        # extra, optional key: `rank = device` takes the extraction order from a stable sort on the GPU instead of numpy's
        # unstable argsort on the host — NOT the reference's tie order (include/icl_b200.h, icl_rank_device); default numpy
        self.rank = str(getattr(self, 'rank', 'numpy')).strip().lower()
        if self.rank not in ('numpy', 'device'):
            raise ValueError(f"rank = {self.rank}: expected 'numpy' (reference tie order) or 'device'")
        self.timing = {}

    # ---- ingest: wgdi/run_colliearity.py:88-121 -------------------------------------------------
    def load_points(self):
        """Read lens/gff/blast, filter, grade and group: returns a dict with the graded points
        (pair_id, loc1, loc2, grading, grade_class), the pair names in output order and the gene tables.
        `grade_class` is the rank class behind every grading (0 red / 1 blue / 2 rest, :64-86 resp. the bins of :48-53):
        a sweep over `grading` keeps the classes and swaps the values (batch.sweep_jobs)."""
        lens1 = base.newlens(self.lens1, 'order')
        lens2 = base.newlens(self.lens2, 'order')
        same = self.gff2 == self.gff1 and self.lens2 == self.lens1       # self comparison: parse the tables once
        gff1 = base.newgff(self.gff1, usecols=range(6))
        gff1 = gff1[gff1['chr'].isin(lens1.index)]        # :96-97
        gff2 = gff1 if same else base.newgff(self.gff2, usecols=range(6))
        if not same:
            gff2 = gff2[gff2['chr'].isin(lens2.index)]
        mode = self.comparison.lower()                    # :108-111
        if mode not in ('genomes', 'chromosomes'):
            raise ValueError(f"comparison = {self.comparison}: expected 'genomes' or 'chromosomes'")
        native = self.native and not _host.is_compressed(self.blast)       # compressed tables: pandas decompresses (base.py:174)
        if native and self.fused and mode == 'genomes':
            return self._load_points_fused(gff1, gff2, same)
        if native:
            ids1 = _host.pack_strings(gff1.index)
            ids2 = ids1 if same else _host.pack_strings(gff2.index)
            hits = _host.read_m8(self.blast, self.blast_reverse == 'true', int(self.score), float(self.evalue), ids1, ids2)
            g1, g2 = hits['g1'], hits['g2']
            bits = hits['bits']
        else:
            blast = base.newblast(self.blast, int(self.score), float(self.evalue), gff1, gff2, self.blast_reverse)
            g1, g2, bits = blast['g1'].to_numpy(), blast['g2'].to_numpy(), blast[11].to_numpy()
        t1 = _gene_table(gff1)
        t2 = t1 if same else _gene_table(gff2)
        if mode == 'chromosomes':                         # :42-57, every sort stable
            name_rank = np.empty(len(gff1), np.int64)
            name_rank[np.argsort(t1['name'], kind='stable')] = np.arange(len(gff1))
            grade, order = grade_hits_chromosomes(name_rank, g1, g2, t1['chr_code'][g1] * len(t2['chr_names']) + t2['chr_code'][g2],
                                                  bits, int(self.multiple), int(self.repeat_number), self.grading, classes=True)
            g1, g2, grade = g1[order], g2[order], grade[order]
        elif self.native:
            grade = grade_hits_native(g1, bits, len(gff1), int(self.multiple), int(self.repeat_number), [1, 2, 3])
        else:
            grade = grade_hits(g1, bits, int(self.multiple), int(self.repeat_number), [1, 2, 3])
        ok = grade > 0                                    # :86 / :57
        g1, g2, cls = g1[ok], g2[ok], grade[ok] - 1
        grade = np.asarray(self.grading, np.int64)[cls]
        print(f"The filtered homologous gene pairs are {len(g1)}.\n")      # :112
        if len(g1) < 1:
            print("Stopped!\n\nIt may be that the id1 and id2 in the BLAST file do not match with (gff1, lens1) and (gff2, lens2).")
            sys.exit(1)                                   # :113-115
        pid, names = pair_ids(t1['chr_names'], t2['chr_names'], t1['chr_code'][g1], t2['chr_code'][g2])
        return dict(pair_id=pid, loc1=t1['order'][g1].astype(np.int32), loc2=t2['order'][g2].astype(np.int32),
                    grading=grade.astype(np.int32), grade_class=cls.astype(np.int32), pair_names=names, genes1=t1, genes2=t2,
                    bounds=_point_bounds(len(names), t1, t2, self.grading))

    def _load_points_fused(self, gff1, gff2, same):
        """`load_points` for `comparison = genomes` through libicl_host.so's fused ingest (include/icl_host.h,
        icl_host_ingest_*): parse, filters, duplicates, per-query grading and the point table are native and threaded;
        Python sees the gene tables, the few queries with boundary ties (numpy ranks them, hazard H2) and the names of
        the chromosome pairs."""
        from concurrent.futures import ThreadPoolExecutor
        ids1 = _host.pack_strings(gff1.index)
        ids2 = ids1 if same else _host.pack_strings(gff2.index)
        with ThreadPoolExecutor(1) as side:               # the gene tables are built while the library parses the m8 file
            tables = side.submit(lambda: (_gene_table(gff1), None if same else _gene_table(gff2)))
            ing = _host.Ingest(self.blast, self.blast_reverse == 'true', int(self.score), float(self.evalue), ids1, ids2,
                               int(self.multiple), int(self.repeat_number))
            try:
                t1, t2 = tables.result()
            except BaseException:
                ing.close()
                raise
        t2 = t1 if same else t2
        with ing:
            off, bits = ing.ambiguous()
            ing.set_classes(classes_of_ambiguous(off, bits, int(self.multiple), int(self.repeat_number)))
            nc1, nc2 = len(t1['chr_names']), len(t2['chr_names'])
            present, n = ing.count(t1['chr_code'], t2['chr_code'], nc1, nc2)
            print(f"The filtered homologous gene pairs are {n}.\n")       # :112
            if n < 1:
                print("Stopped!\n\nIt may be that the id1 and id2 in the BLAST file do not match with (gff1, lens1) and (gff2, lens2).")
                sys.exit(1)                               # :113-115
            table, names = pair_table(t1['chr_names'], t2['chr_names'], present)
            pts = ing.points(t1['order'], t2['order'], nc2, table, self.grading)
        pts.update(pair_names=names, genes1=t1, genes2=t2)
        pts['bounds'] = _point_bounds(len(names), t1, t2, self.grading)
        return pts

    # ---- wgdi/run_colliearity.py:88-156 ----------------------------------------------------------
    def _engine_in_background(self):
        """The device side of a run needs nothing from the input files to get ready: the CUDA context of a fresh process
        (~1 s), the engine's streams and the start-up check of the device rank against the running numpy are started here and
        proceed while the host parses and grades (the native ingest and the CUDA calls both run without the GIL)."""
        from concurrent.futures import ThreadPoolExecutor
        if len(self.devices) != 1:
            # sharded run: every GPU's context is created now, side by side (chain_on_devices builds its engines later,
            # from the cached blocks and on live contexts); a failure here is reported by the run itself
            def touch(d):
                try:
                    IclEngine(d).close()
                except Exception:
                    pass
            pool = ThreadPoolExecutor(len(set(self.devices)))
            futs = [pool.submit(touch, d) for d in sorted(set(self.devices))]
            pool.shutdown(wait=False)
            self._warm_devices = futs
            return None

        def make():
            eng = IclEngine(self.devices[0])
            if self.rank != 'device' and eng.rank_mode == 'auto':
                eng.numpy_sort_kind()
            return eng
        pool = ThreadPoolExecutor(1)
        fut = pool.submit(make)
        pool.shutdown(wait=False)
        return fut

    def run(self):
        t0 = time.perf_counter()
        warm = self._engine_in_background()
        try:
            pts = self.load_points()
        except BaseException:                             # sys.exit(1) of :113-115 included: no engine is left behind
            if warm is not None:
                try:
                    warm.result().close()
                except Exception:
                    pass
            raise
        t1 = time.perf_counter()
        params = ChainParams.from_options(self.options)
        for c1, c2 in pts['pair_names']:
            print(f'Running {c1} vs {c2}')                # :162
        eng = None
        try:
            if len(self.devices) > 1:
                if self.rank == 'device':
                    raise ValueError("rank = device is offered for a single device only")
                from .shard import chain_on_devices
                for f in getattr(self, '_warm_devices', ()):
                    f.result()
                blocks = chain_on_devices(pts, len(pts['pair_names']), params, self.devices, min_members=self.over_gap)
            else:
                eng = warm.result()
                eng.fields = "writer"     # only what the file is written from comes back: 32 B per block, 8 B per gene pair
                # the points cross PCIe in the narrowest element types that hold them (7 instead of 16 bytes per point)
                up = compact_points(pts['pair_id'], pts['loc1'], pts['loc2'], pts['grading'], pts.get('bounds'))
                # blocks are views of the engine's page-locked staging: written out before the engine is closed
                if self.rank == 'device':
                    eng.upload(*up, len(pts['pair_names']))
                    blocks = eng.run_resident_device_rank(params, min_members=self.over_gap, copy=False)
                else:
                    blocks = eng.chain(*up, len(pts['pair_names']), params, min_members=self.over_gap, copy=False)
                self.timing['device_ms'] = eng.timings()
                self.timing['host_ms'] = dict(eng.host_ms)
            t2 = time.perf_counter()
            if self.native:
                write_blocks_native(self.savefile, blocks, pts['pair_names'], pts['genes1'], pts['genes2'], params.pvalue,
                                    self.over_gap)
            else:
                text = format_blocks(blocks, pts['pair_names'], pts['genes1'], pts['genes2'], params.pvalue, self.over_gap)
                with open(self.savefile, 'w') as fout:
                    fout.write(text)
            del blocks
        finally:
            if eng is not None:
                eng.close()
        t3 = time.perf_counter()
        self.timing.update(ingest_s=t1 - t0, chain_s=t2 - t1, write_s=t3 - t2)
        sys.exit(0)                                       # :156


def _point_bounds(n_pairs, genes1, genes2, grading):
    """(lowest, highest) value the four point arrays can hold, known without looking at them (`compact_points`)."""
    def span(order):
        return (int(order.min()), int(order.max())) if len(order) else (0, 0)
    return ((0, max(int(n_pairs) - 1, 0)), span(genes1['order']), span(genes2['order']),
            (min(int(g) for g in grading), max(int(g) for g in grading)))


def _gene_table(gff):
    """Per-genome lookup tables: chromosome code per gene row, (chromosome, order) -> name / strand."""
    chr_names, chr_code = np.unique(gff['chr'].to_numpy().astype(str), return_inverse=True)
    order = gff['order'].to_numpy().astype(np.int64)
    maxo = int(order.max()) + 1 if len(order) else 1
    row_of = np.full(len(chr_names) * maxo, -1, np.int64)
    # like Series.map on an index with duplicates would, the first gene of an order wins
    key = chr_code.astype(np.int64) * maxo + order
    row_of[key[::-1]] = np.arange(len(order))[::-1]
    strand = gff['strand'].to_numpy().astype(str)
    return dict(chr_names=list(chr_names), chr_code=chr_code.astype(np.int64), order=order, maxo=maxo, row_of=row_of,
                name=gff.index.to_numpy().astype(str), strand=strand)


def _native_gene_tables(genes1, genes2):
    """Tables for icl_host_write_blocks: packed names and strand codes over a vocabulary shared by both genomes."""
    vocab = {v: i for i, v in enumerate(sorted(set(genes1['strand'].tolist()) | set(genes2['strand'].tolist())))}
    for g in (genes1, genes2):
        if 'name_bytes' not in g:
            g['name_bytes'], g['name_off'] = _host.pack_strings(g['name'])
            g['strand_code'] = np.array([vocab[v] for v in g['strand']], np.int32)
            g['row_of'] = np.ascontiguousarray(g['row_of'], np.int64)


def write_blocks_native(path, blocks, pair_names, genes1, genes2, pvalue, over_gap):
    """Same file as `format_blocks`, assembled and written by libicl_host.so."""
    p = blocks.p_value()                                  # numpy rounding, collinearity.py:153
    emit = np.flatnonzero(~(p > pvalue) & (blocks.m >= over_gap))
    _native_gene_tables(genes1, genes2)
    c1code = {name: i for i, name in enumerate(genes1['chr_names'])}
    c2code = {name: i for i, name in enumerate(genes2['chr_names'])}
    pc1 = np.array([c1code[a] for a, _ in pair_names], np.int32)
    pc2 = np.array([c2code[b] for _, b in pair_names], np.int32)
    beg, end = blocks.member_off[emit], blocks.member_off[emit + 1]
    # p values are round(a, 4) (collinearity.py:153): a few thousand distinct numbers, printed by Python itself
    pu, pidx = np.unique(p[emit], return_inverse=True)
    pair = np.asarray(blocks.pair)[emit]
    _host.write_collinearity(path, np.asarray(blocks.score)[emit], [f"{v}" for v in pu], pidx,
                             [f"{a}&{b}" for a, b in pair_names], pair, beg, end, blocks.loc1, blocks.loc2,
                             pc1[pair], pc2[pair], genes1, genes2)


def format_blocks(blocks, pair_names, genes1, genes2, pvalue, over_gap):
    """The collinearity file body (wgdi/run_colliearity.py:143-155, 174-199)."""
    p = blocks.p_value()                                  # round(a, 4), collinearity.py:153
    emit = np.flatnonzero(~(p > pvalue) & (blocks.m >= over_gap))          # collinearity.py:55-56, run_colliearity.py:178
    c1code = {name: i for i, name in enumerate(genes1['chr_names'])}
    c2code = {name: i for i, name in enumerate(genes2['chr_names'])}
    out = []
    num = 1
    for b in emit:
        a0, a1 = blocks.member_off[b], blocks.member_off[b + 1]
        x, y = blocks.loc1[a0:a1], blocks.loc2[a0:a1]
        chr1, chr2 = pair_names[blocks.pair[b]]
        mark = 'plus' if y[0] < y[1] else 'minus'         # :194-195
        tail = f" score={int(blocks.score[b])} pvalue={p[b]} N={a1 - a0} {chr1}&{chr2} {mark}"
        out.append(f"# Alignment {num}:" + tail.split(':')[0])             # :147-151 re-cuts the header at ':'
        num += 1
        r1 = genes1['row_of'][c1code[chr1] * genes1['maxo'] + x]
        r2 = genes2['row_of'][c2code[chr2] * genes2['maxo'] + y]
        n1, n2 = genes1['name'][r1], genes2['name'][r2]
        s = np.where(genes1['strand'][r1] == genes2['strand'][r2], '1', '-1')        # :185-187
        out.extend(f"{a} {i} {c} {j} {k}" for a, i, c, j, k in zip(n1, x.tolist(), n2, y.tolist(), s))
    return "\n".join(out) + "\n" if out else ""
