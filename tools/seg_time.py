"""Rank-kernel time for subsets of a config's chromosome pairs (largest pair alone, a median pair, the 8 / 74 largest):
how much of the rank phase is one segment's critical path.   [ICL_B200_XS_WARPS=n] python tools/seg_time.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_points, params_for
from wgdi_b200.engine import IclEngine
pts = make_points("C2")
prm = params_for(pts["cfg"])
cnt = np.bincount(pts["pair_id"], minlength=pts["n_pairs"])
order = np.argsort(-cnt)
eng = IclEngine(0)
kind = eng.numpy_sort_kind()
for label, pairs in (("largest pair", order[:1]), ("median pair", order[len(order) // 2:len(order) // 2 + 1]), ("8 largest", order[:8]), ("74 largest", order[:74])):
    m = np.isin(pts["pair_id"], pairs)
    remap = np.full(pts["n_pairs"], -1, np.int32); remap[np.sort(pairs)] = np.arange(len(pairs), dtype=np.int32)
    eng.upload(remap[pts["pair_id"][m]], pts["loc1"][m], pts["loc2"][m], pts["grading"][m], len(pairs))
    ts = []
    for _ in range(4):
        eng.run_resident_device_rank(prm, min_members=5, mode=kind)
        ts.append(eng.timings()["keys_d2h"])
    print(os.environ.get("ICL_B200_XS_WARPS"), label, "points", int(m.sum()), "rank ms", round(min(ts), 4), flush=True)
eng.close()
