"""Drop-in for the reference's DP engine class: `collinearity(options, points).run()`
(wgdi/collinearity.py:5-70), one chromosome pair per call, executed on the GPU.

`points` is the frame `single_pool` builds (wgdi/run_colliearity.py:164-166): columns
loc1, loc2, grading, sorted by (loc1, loc2), unique index.  `run()` returns the reference's
`[[DataFrame[loc1, loc2] sorted by loc1, p_value, score], ...]`, forward blocks first,
extraction order preserved, blocks of 3-4 pairs included, p filtered with `pvalue`.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from .engine import ChainParams, IclEngine

_ENGINES = {}


def _engine(device):
    if device not in _ENGINES:
        _ENGINES[device] = IclEngine(device)
    return _ENGINES[device]


class collinearity:
    def __init__(self, options, points, device=0):
        self.options = options
        self.points = points
        self.params = ChainParams.from_options(options)
        self.device = device

    def run(self):
        pts = self.points
        n = len(pts)
        if n == 0:
            return []
        eng = _engine(self.device)
        loc1 = pts['loc1'].to_numpy().astype(np.int32)
        loc2 = pts['loc2'].to_numpy().astype(np.int32)
        blocks = eng.chain(np.zeros(n, np.int32), loc1, loc2, pts['grading'].to_numpy().astype(np.int32), 1, self.params)
        p = blocks.p_value()
        labels = pts.index.to_numpy()
        data = []
        for b in range(len(blocks)):
            if p[b] > self.params.pvalue:                       # collinearity.py:55-56
                continue
            a0, a1 = blocks.member_off[b], blocks.member_off[b + 1]
            frame = pd.DataFrame({'loc1': blocks.loc1[a0:a1].astype(np.int64), 'loc2': blocks.loc2[a0:a1].astype(np.int64)},
                                 index=labels[blocks.member_idx[a0:a1]])
            data.append([frame, p[b], int(blocks.score[b])])
        return data
