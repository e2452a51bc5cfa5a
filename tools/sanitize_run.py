"""Small pass of the whole hot path for compute-sanitizer (memcheck / racecheck): tiny + cross datasets, all DP kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wgdi_b200 import synth
from wgdi_b200.engine import ChainParams, IclEngine

for name in ("tiny", "cross"):
    ds = synth.make_dataset(name)
    pts = synth.graded_points(ds)
    prm = ChainParams(mg1=ds.cfg.mg[0], mg2=ds.cfg.mg[1], grading=tuple(ds.cfg.grading))
    eng = IclEngine(0)
    blocks = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
    print(name, os.environ.get("ICL_B200_CHAIN", "ring"), os.environ.get("ICL_B200_PVALUE", "auto"), len(pts["loc1"]), "points", len(blocks), "blocks", eng.counters()["chain_kernel"])
    eng.close()
