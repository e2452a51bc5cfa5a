"""BASELINE.json config 5: a batch of 28 pairwise genome comparisons (8 species) x the parameter sweep
grading in {50,40,25; 50,30,25} x mg in {25,25; 40,40; 50,50} x pvalue in {0.05, 0.2, 1} on one or several GPUs.

    python tools/config5.py [--scale S] [--gpus 0,1,...] [--check K]

Species are synthetic (wgdi_b200/synth.py: 25-45 k genes, 7-19 chromosomes at scale 1); each of the 28 comparisons
is generated as its own cross-species dataset.  Per comparison the sweep needs 6 device jobs (the 3 pvalue variants of
a (grading, mg) share one: `pvalue` only filters the output, wgdi/collinearity.py:55-56) -> 168 jobs, 504 results.
Jobs are dealt to the GPUs comparison by comparison (longest first) and pipelined through two engines per GPU
(wgdi_b200/batch.py).  --check K compares the blocks of K jobs with the C oracle (same numpy tie order).
"""
import argparse
import itertools
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from wgdi_b200 import synth
from wgdi_b200.batch import BatchRunner, sweep_jobs
from wgdi_b200.engine import ChainParams
from wgdi_b200.shard import lpt_assign

GRADINGS, MGS, PVALUES, SPECIES = synth.C5_GRADINGS, synth.C5_MGS, synth.C5_PVALUES, synth.C5_SPECIES


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--gpus", default="0")
    ap.add_argument("--check", type=int, default=3)
    args = ap.parse_args()
    gpus = [int(g) for g in args.gpus.split(",")]
    t0 = time.perf_counter()
    comps, cfgs = [], []
    for ci, (a, b) in enumerate(itertools.combinations(range(len(SPECIES)), 2)):
        cfg = synth.config5_comparison(ci)
        cfgs.append(cfg)
        comps.append(synth.graded_points(synth.make_dataset(cfg, scale=args.scale)))
    n_hits = [len(c["loc1"]) for c in comps]
    print(f"{len(comps)} comparisons, {sum(n_hits)} graded hits in total ({min(n_hits)}..{max(n_hits)}), "
          f"generated in {time.perf_counter() - t0:.1f} s", flush=True)
    owner = lpt_assign(n_hits, len(gpus))
    results = {}
    lock = threading.Lock()

    def on_gpu(gi):
        runner = BatchRunner(gpus[gi], depth=4, sort_threads=max(1, (os.cpu_count() or 1) // len(gpus)))
        try:
            mine = [c for c in np.argsort(-np.asarray(n_hits)) if owner[c] == gi]
            jobs, keys = [], []
            for c in mine:
                j, k = sweep_jobs(comps[c], GRADINGS, MGS, PVALUES)
                jobs += j
                keys += [(int(c),) + kk for kk in k]
            # every result of the sweep: per pvalue the number of blocks the file writer would keep
            def consume(i, blocks):
                p = blocks.p_value()
                keep5 = blocks.m >= 5
                out = {pv: int((~(p > pv) & keep5).sum()) for pv in keys[i][3]}
                want = keys[i][:3] in check_keys
                return out, (blocks.pair.copy(), blocks.dir.copy(), blocks.score.copy(), blocks.p_raw.copy(), blocks.member_off.copy(),
                             blocks.loc1.copy(), blocks.loc2.copy()) if want else None
            outs = runner.run(jobs, min_members=1, consume=consume)
            with lock:
                for k, o in zip(keys, outs):
                    results[k[:3]] = o
        finally:
            runner.close()

    rng = np.random.default_rng(5)
    all_keys = [(c, g, mg) for c in range(len(comps)) for g in GRADINGS for mg in MGS]
    check_keys = set(all_keys[i] for i in rng.choice(len(all_keys), min(args.check, len(all_keys)), replace=False))
    t1 = time.perf_counter()
    th = [threading.Thread(target=on_gpu, args=(gi,)) for gi in range(len(gpus))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t1
    assert len(results) == len(all_keys)
    n_results = sum(len(o[0]) for o in results.values())
    jobs_hits = sum(n_hits[c] for c, _, _ in all_keys)
    print(f"{len(all_keys)} device jobs = {n_results} sweep results on {len(gpus)} GPU(s) in {dt:.2f} s: "
          f"{jobs_hits / dt / 1e6:.1f} M hits/s, {dt / len(all_keys) * 1e3:.2f} ms per job", flush=True)
    if check_keys:
        from conftest import oracle_blocks
        for key in sorted(check_keys):
            c, g, mg = key
            prm = ChainParams(mg1=mg[0], mg2=mg[1], grading=g)
            # the oracle's input is graded from scratch with THIS job's grading (run_colliearity.py:64-86), not looked up
            # from the classes the device path uses
            own = synth.graded_points(synth.make_dataset(cfgs[c], scale=args.scale), grading=g)
            exp = oracle_blocks(own, prm)
            pair, d, score, p_raw, moff, l1, l2 = results[key][1]
            assert len(exp) == len(pair), (key, len(exp), len(pair))
            for b, (ep, ed, el1, el2, ea, esc) in enumerate(exp):
                assert pair[b] == ep and d[b] == ed and score[b] == esc and p_raw[b] == ea
                assert (l1[moff[b]:moff[b + 1]] == el1).all() and (l2[moff[b]:moff[b + 1]] == el2).all()
            print(f"checked comparison {c} grading {g} mg {mg}: {len(exp)} blocks identical to the oracle", flush=True)


if __name__ == "__main__":
    main()
