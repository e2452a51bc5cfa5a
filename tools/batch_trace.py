"""Timeline of the pipelined batch (wgdi_b200/batch.py): per job and engine, when each phase started and ended."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from bench import make_points, params_for
from wgdi_b200.batch import BatchRunner
from wgdi_b200.engine import IclEngine

pts = make_points(sys.argv[1] if len(sys.argv) > 1 else "C2")
prm = params_for(pts["cfg"])
njobs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pinned = "--pinned" in sys.argv
T0 = time.perf_counter()
log = []
orig_upload, orig_stream, orig_extract = IclEngine.upload, IclEngine.dp_rank_streamed, IclEngine.extract


def wrap(name, fn):
    def f(self, *a, **k):
        t = time.perf_counter()
        r = fn(self, *a, **k)
        log.append((id(self) % 1000, name, (t - T0) * 1e3, (time.perf_counter() - T0) * 1e3, dict(self.host_ms) if name == "dp+rank" else None))
        return r
    return f


IclEngine.upload = wrap("upload", orig_upload)
IclEngine.dp_rank_streamed = wrap("dp+rank", orig_stream)
IclEngine.extract = wrap("extract", orig_extract)
_lib_extract = None


def patch_lib(eng):
    # time the C call itself (inside extract)
    orig = eng.lib.icl_extract

    def timed(*a):
        t = time.perf_counter()
        r = orig(*a)
        log.append((0, "  icl_extract", (t - T0) * 1e3, (time.perf_counter() - T0) * 1e3, None))
        return r
    return timed
runner = BatchRunner(0, depth=int(os.environ.get("DEPTH", "2")), sort_threads=int(os.environ["SORT_THREADS"]) if os.environ.get("SORT_THREADS") else None)
_real = runner.engines[0].lib.icl_extract
class _Lib:
    def __init__(self, lib): self._lib = lib
    def __getattr__(self, k):
        if k == "icl_extract":
            def timed(*a):
                t = time.perf_counter()
                r = self._lib.icl_extract(*a)
                log.append((0, "  C.extract", (t - T0) * 1e3, (time.perf_counter() - T0) * 1e3, None))
                return r
            return timed
        return getattr(self._lib, k)
for e in runner.engines:
    e.lib = _Lib(e.lib)
if pinned:
    e0 = runner.engines[0]
    host = {}
    for k in ("pair_id", "loc1", "loc2", "grading"):
        buf = e0._buf("in_" + k, len(pts[k]), np.int32)
        buf[:] = pts[k]
        host[k] = buf
    host["n_pairs"] = pts["n_pairs"]
    pts = host
jobs = [(dict(pts), prm) for _ in range(njobs)]
runner.run(jobs[:2], min_members=5, consume=(lambda i, b: len(b)) if '--views' in sys.argv else None)
log.clear()
T0 = time.perf_counter()
t = time.perf_counter()
runner.run(jobs, min_members=5, consume=(lambda i, b: len(b)) if '--views' in sys.argv else None)
dt = time.perf_counter() - t
runner.close()
print(f"{njobs} jobs in {dt * 1e3:.2f} ms = {dt * 1e3 / njobs:.2f} ms per comparison")
for e in sorted(log, key=lambda x: x[2]):
    print(f"eng {e[0]:4d} {e[1]:8s} {e[2]:8.2f} -> {e[3]:8.2f}  ({e[3] - e[2]:6.2f})", e[4] or "")
