"""Data half of `wgdi -d`: which BLAST hits the dotplot draws and in which colour.

The reference's `dotplot.run()` (wgdi/dotplot.py:58-123) reads lens / gff / blast like `-icl`, colours every hit by its
per-query rank (`pair_positon`, :31-56) and hands the coloured points to matplotlib.  The plotting is out of scope
(matplotlib is not in the image); the rank-to-colour classes are the same per-query rank as the
`-icl` grading with a wider blue class, so they are offered here as a data function on the same ingest path.
"""
from __future__ import annotations

import numpy as np

from . import _host, base
from .ingest import COLOR_NAMES, color_hits


class dotplot():
    def __init__(self, options):
        # defaults of wgdi/dotplot.py:11-21 (the plotting ones are accepted and unused)
        self.multiple = 1
        self.score = 100
        self.evalue = 1e-5
        self.repeat_number = 20
        self.position = 'order'
        self.blast_reverse = False
        for k, v in options:
            setattr(self, str(k), v)
        self.native = str(getattr(self, 'ingest', 'native')).lower() != 'pandas' and _host.available()

    def colored_hits(self):
        """dict(row, g1, g2, color, color_name): the hits `pair_positon` keeps, in BLAST row order.
        row = row label of the m8 table, g1 / g2 = gene row in gff1 / gff2 (lens-filtered, as gene_location does at
        base.py:222), color = 1 red / 2 blue / 3 gray."""
        lens1 = base.newlens(self.lens1, self.position)
        lens2 = base.newlens(self.lens2, self.position)
        gff1 = base.newgff(self.gff1)
        gff1 = gff1[gff1['chr'].isin(lens1.index)]
        same = self.gff2 == self.gff1 and self.lens2 == self.lens1
        gff2 = gff1 if same else base.newgff(self.gff2)
        if not same:
            gff2 = gff2[gff2['chr'].isin(lens2.index)]
        # `base.str_to_bool(self.blast_reverse)` at dotplot.py:29 discards its result, so the STRING of the conf reaches
        # base.newblast (:108-109) and 'true' swaps the id columns — unlike -bi, which stores the bool
        reverse = self.blast_reverse == 'true'
        if self.native and not _host.is_compressed(self.blast):
            ids1 = _host.pack_strings(gff1.index)
            ids2 = ids1 if same else _host.pack_strings(gff2.index)
            hits = _host.read_m8(self.blast, reverse, int(self.score), float(self.evalue), ids1, ids2)
            row, g1, g2, bits = hits['row'], hits['g1'], hits['g2'], hits['bits']
        else:
            blast = base.newblast(self.blast, int(self.score), float(self.evalue), gff1, gff2, 'true' if reverse else 'false')
            row, g1, g2, bits = blast.index.to_numpy(), blast['g1'].to_numpy(), blast['g2'].to_numpy(), blast[11].to_numpy()
        color = color_hits(g1, bits, int(self.multiple), int(self.repeat_number))
        keep = color > 0                                   # dotplot.py:56
        return dict(row=np.asarray(row)[keep], g1=np.asarray(g1)[keep], g2=np.asarray(g2)[keep], color=color[keep],
                    color_name=np.asarray(COLOR_NAMES, dtype=object)[color[keep]],
                    order1=gff1['order'].to_numpy()[np.asarray(g1)[keep]], order2=gff2['order'].to_numpy()[np.asarray(g2)[keep]])
