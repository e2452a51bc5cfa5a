"""Short driver for ncu captures: one comparison (default C2), a few passes of the hot path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import make_points, params_for
from wgdi_b200.engine import IclEngine

workload = sys.argv[1] if len(sys.argv) > 1 else "C2"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pts = make_points(workload)
prm = params_for(pts["cfg"])
eng = IclEngine(0)
eng.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"])
for _ in range(passes):
    blocks = eng.run_resident(prm, min_members=5)
print("points", len(pts["loc1"]), "blocks", len(blocks), eng.timings(), eng.counters())
eng.close()
