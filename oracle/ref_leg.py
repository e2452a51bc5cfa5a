"""CPU baseline leg that runs the REAL reference: `wgdi.collinearity.collinearity(options, points).run()`
(wgdi/collinearity.py:42-70, the inner seam `single_pool` calls per chromosome pair, run_colliearity.py:168-170) from the
unmodified copy under oracle/_ref/ (oracle/make_ref.py), one chromosome pair per worker process.

The reference needs 6-8 ms per point per core, hours for a BASELINE config, so only a
stratified sample of pairs is run and the whole-workload figure is an EXTRAPOLATION, labelled as such, with the
per-pair model BASELINE.md fitted (t = a*n + b*n^2, re-fitted here on the sample).  Test infrastructure / benchmark only.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "wgdi", "collinearity.py"))


def run_pair(job):
    """Worker: one chromosome pair through the reference.  job = (loc1, loc2, grading, options) with the points in
    (loc1, loc2) order.  Returns (seconds, [(loc1[], loc2[], p, score)] in the reference's emission order)."""
    import io
    import contextlib
    import pandas as pd
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import wgdi.collinearity as ref                       # numpy + pandas only (collinearity.py:1-2)
    loc1, loc2, grading, options = job
    points = pd.DataFrame({"loc1": np.asarray(loc1, np.int64), "loc2": np.asarray(loc2, np.int64),
                           "grading": np.asarray(grading, np.int64)})
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        data = ref.collinearity(options, points).run()
    dt = time.perf_counter() - t0
    out = [(blk["loc1"].to_numpy().astype(np.int64), blk["loc2"].to_numpy().astype(np.int64), float(p), int(sc))
           for blk, p, sc in data]
    return dt, out


def stratified_pairs(sizes, k=8, max_points=None, seed=7):
    """Largest, median, smallest and k - 3 random pairs among those with 3 <= points <= max_points."""
    sizes = np.asarray(sizes)
    ok = np.flatnonzero((sizes >= 3) & ((sizes <= max_points) if max_points else True))
    if len(ok) < 3:                                         # nothing that small: the k smallest pairs instead
        ok = np.flatnonzero(sizes >= 3)
        ok = ok[np.argsort(sizes[ok], kind="stable")[:max(1, k)]]
    if len(ok) == 0:
        return []
    by = ok[np.argsort(sizes[ok], kind="stable")]
    pick = [int(by[-1]), int(by[len(by) // 2]), int(by[0])]
    rest = [int(p) for p in np.random.default_rng(seed).permutation(by) if int(p) not in pick]
    return (pick + rest)[:max(1, k)]


def time_sample(pts, options, k=8, max_points=None, workers=None):
    """Run the reference on a stratified sample of the chromosome pairs of `pts` (dict pair_id, loc1, loc2, grading,
    n_pairs), one pair per process.  Returns a dict: per-pair sizes and seconds, the re-fitted model, the extrapolated
    whole-workload time on `cores` cores, and the reference's blocks per sampled pair (for parity checks)."""
    import pickle
    import subprocess
    import tempfile
    sizes = np.bincount(pts["pair_id"], minlength=pts["n_pairs"])
    pairs = stratified_pairs(sizes, k, max_points)
    jobs = []
    for p in pairs:
        sel = np.flatnonzero(pts["pair_id"] == p)
        o = sel[np.lexsort((pts["loc2"][sel], pts["loc1"][sel]))]            # run_colliearity.py:164-166, ties in file order
        jobs.append((pts["loc1"][o], pts["loc2"][o], pts["grading"][o], list(options)))
    cores = os.cpu_count() or 1
    workers = min(workers or cores, len(jobs)) or 1
    # one fresh interpreter per pair (`python oracle/ref_leg.py --pair job out`): nothing of the parent — CUDA context,
    # torch, the product's libraries — is inherited, `workers` of them at a time
    tmp = tempfile.mkdtemp(prefix="icl_ref_leg_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    t0 = time.perf_counter()
    try:
        for i, job in enumerate(jobs):
            with open(os.path.join(tmp, f"job{i}.pkl"), "wb") as f:
                pickle.dump(job, f)
        running, res, nxt = {}, [None] * len(jobs), 0
        while nxt < len(jobs) or running:
            while nxt < len(jobs) and len(running) < workers:
                running[nxt] = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--pair",
                                                 os.path.join(tmp, f"job{nxt}.pkl"), os.path.join(tmp, f"out{nxt}.pkl")])
                nxt += 1
            for i, pr in list(running.items()):
                if pr.poll() is not None:
                    if pr.returncode != 0:
                        raise RuntimeError(f"reference worker for pair {pairs[i]} exited with {pr.returncode}")
                    with open(os.path.join(tmp, f"out{i}.pkl"), "rb") as f:
                        res[i] = pickle.load(f)
                    del running[i]
            time.sleep(0.02)
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    wall = time.perf_counter() - t0
    n = np.array([len(j[0]) for j in jobs], np.float64)
    t = np.array([r[0] for r in res], np.float64)
    # t = a*n + b*n^2 (least squares, both >= 0), BASELINE.md section 2's model re-fitted on this box
    A = np.stack([n, n * n], 1)
    coef, *_ = np.linalg.lstsq(A, t, rcond=None)
    if coef[1] < 0 or coef[0] < 0:
        coef = np.array([float(t.sum() / n.sum()), 0.0])
    per_pair = coef[0] * sizes + coef[1] * sizes.astype(np.float64) ** 2
    # the reference's pool: static contiguous chunks of ceil(P / process) pairs, one per process (run_colliearity.py:123-132);
    # the run lasts as long as its slowest chunk
    P = len(sizes)
    chunk = -(-P // cores)
    chunks = [per_pair[i:i + chunk].sum() for i in range(0, P, chunk)]
    total_cpu = float(per_pair.sum())
    est_wall = float(max(chunks)) if chunks else 0.0
    return {"pairs": [int(p) for p in pairs], "points": [int(x) for x in n], "seconds": [float(x) for x in t],
            "sample_wall_s": wall, "workers": workers, "cores": cores,
            "model": {"a_s_per_point": float(coef[0]), "b_s_per_point2": float(coef[1])},
            "measured_points_per_s_per_core": float(n.sum() / t.sum()) if t.sum() > 0 else None,
            "extrapolated_cpu_s": total_cpu, "extrapolated_wall_s_all_cores": est_wall,
            "extrapolated_points_per_s_all_cores": float(sizes.sum() / est_wall) if est_wall > 0 else None,
            "blocks": {int(p): r[1] for p, r in zip(pairs, res)}}


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--pair":
        import pickle
        with open(sys.argv[2], "rb") as f:
            job = pickle.load(f)
        with open(sys.argv[3], "wb") as f:
            pickle.dump(run_pair(job), f)
