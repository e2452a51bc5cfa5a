"""Ingest of one conf through the three host paths (fused native / native steps / pandas): times and equality of the points.

    python tools/ingest_compare.py <dir with icl.conf> [--pandas]

`synth.write_dataset(synth.make_dataset("C4"), dir)` writes such a directory.  ICL_HOST_TIMING=1 adds the library's phase times."""
import contextlib
import io
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wgdi_b200 import base                                        # noqa: E402
from wgdi_b200.run_colliearity import mycollinearity              # noqa: E402


def main():
    conf = os.path.join(sys.argv[1], "icl.conf")
    modes = ["native", "native-steps", "native", "native-steps"] + (["pandas"] if "--pandas" in sys.argv else [])
    res = {}
    for mode in modes:
        with contextlib.redirect_stdout(io.StringIO()):
            m = mycollinearity(list(base.load_conf(conf, "collinearity")) + [("ingest", mode)])
            t = time.perf_counter()
            pts = m.load_points()
            dt = time.perf_counter() - t
        print(f"{mode:13s} {dt:7.3f} s  {len(pts['loc1'])} points", flush=True)
        res[mode] = pts
    a = res["native"]
    for other, b in res.items():
        for k in ("pair_id", "loc1", "loc2", "grading", "grade_class"):
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), (other, k)
        assert a["pair_names"] == b["pair_names"], other
    print("identical points in all paths")


if __name__ == "__main__":
    main()
