"""Batches of comparisons on one GPU (BASELINE.json config 5: 28 pairwise genome comparisons x a grading / mg /
pvalue sweep).

One comparison leaves the GPU idle while the host runs numpy's argsort (the reference's tie order, engine.py) and
leaves most host cores idle during the device phases.  Comparisons are independent, so `BatchRunner` keeps `depth`
engines (contexts with their own stream and page-locked buffers) on the same GPU and runs the jobs through them from
`depth` host threads: while one job is being ranked — the sort pool takes one job at a time — the next job's upload,
sort, edge enumeration and DP run on the device and the previous job's extraction finishes.  Every job is processed
exactly as `IclEngine.chain` would; results come back in job order.

    job k   : upload  dp ──► rank ───────► extract
    job k+1 :            upload  dp ─ (waits for the pool) ─► rank ───────► extract

The `pvalue` option only filters the output (wgdi/collinearity.py:55-56), so parameter sets that differ only in
`pvalue` share one job (`sweep_jobs`).
"""
from __future__ import annotations

import queue
import threading
from dataclasses import replace

import numpy as np

from .engine import ChainParams, IclEngine


class BatchRunner:
    def __init__(self, device: int = 0, depth: int = 4, sort_threads: int | None = None):
        self.engines = [IclEngine(device, sort_threads) for _ in range(max(1, int(depth)))]

    def close(self):
        for e in self.engines:
            e.close()

    def run(self, jobs, min_members: int = 1, consume=None):
        """jobs: sequence of (points, params) with points = dict(pair_id, loc1, loc2, grading, n_pairs) (host arrays;
        consecutive jobs that pass the SAME dict object reuse the upload) -> list of Blocks in job order.
        A points dict with `grade_class` instead of `grading` (what `sweep_jobs` builds) is uploaded once as rank
        classes and every job chains with ITS params.grading through the device's grading table.
        With `consume(job_index, blocks)` the blocks are handed over as zero-copy views of the engine's page-locked
        staging (valid only during the call — e.g. the file writer) and the list holds what `consume` returned."""
        jobs = list(jobs)
        results = [None] * len(jobs)
        errors = []
        todo = queue.Queue()
        for i in range(len(jobs)):
            todo.put(i)

        def worker(eng):
            uploaded = None
            while not errors:
                try:
                    i = todo.get_nowait()
                except queue.Empty:
                    return
                pts, prm = jobs[i]
                try:
                    by_class = "grading" not in pts
                    if uploaded is not pts:
                        eng.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grade_class" if by_class else "grading"],
                                   pts["n_pairs"])
                        uploaded = pts
                    eng.set_grading_table(prm.grading if by_class else None)
                    if consume is None:
                        results[i] = eng.run_resident(prm, min_members, copy=True)
                    else:
                        results[i] = consume(i, eng.run_resident(prm, min_members, copy=False))
                except Exception as e:                     # noqa: BLE001 — re-raised in the caller's thread
                    errors.append(e)
                    return

        threads = [threading.Thread(target=worker, args=(e,)) for e in self.engines[:max(1, min(len(self.engines), len(jobs)))]]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results


def sweep_jobs(points, gradings, mgs, pvalues, base: ChainParams | None = None):
    """The parameter sweep of config 5 over one comparison: one job per (grading, mg); the pvalue variants of a job
    share its blocks (apply `Blocks.p_value() > pvalue` per variant).  Returns (jobs, [(grading, mg, pvalues)]).

    `grading` enters the reference when the hits are graded (wgdi/run_colliearity.py:30-33,64-86): every point's value
    changes with it, not only the divisor of the p value.  The rank CLASS of a hit does not depend on the values, so
    `points` must carry `grade_class` (0 red / 1 blue / 2 rest, as synth.graded_points and mycollinearity.load_points
    return it); the jobs share one upload of the classes and each chains with its own table."""
    if "grade_class" not in points:
        raise ValueError("sweep_jobs needs points['grade_class']: a grading sweep re-grades every hit")
    base = base or ChainParams()
    by_class = {k: points[k] for k in ("pair_id", "loc1", "loc2", "grade_class", "n_pairs")}
    jobs, keys = [], []
    for g in gradings:
        if len(g) <= int(np.max(points["grade_class"], initial=0)):
            raise ValueError("grading has fewer values than there are rank classes")
        for mg in mgs:
            jobs.append((by_class, replace(base, grading=tuple(int(x) for x in g), mg1=int(mg[0]), mg2=int(mg[1]), pvalue=1.0)))
            keys.append((tuple(g), tuple(mg), tuple(pvalues)))
    return jobs, keys


def regrade(points, grading):
    """The points of a comparison as `-icl` with this `grading` would grade them (host copy; tests and checks)."""
    out = dict(points)
    out["grading"] = np.asarray(grading, np.int32)[points["grade_class"]]
    return out
