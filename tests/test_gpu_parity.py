"""Parity of the CUDA path (through the C ABI) against the reference-generated fixtures and the oracle.
Everything here needs a B200: run with `-m gpu`."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import (PORTABLE_ENV, ROOT, check_blocks_against_golden, check_dp_against_golden, oracle_batch,
                      oracle_blocks, oracle_blocks_stable, stage_e2e)

pytestmark = pytest.mark.gpu


def _compare_with_oracle(blocks, exp):
    assert len(blocks) == len(exp)
    for b, (pair, d, l1, l2, a, sc) in enumerate(exp):
        a0, a1 = blocks.member_off[b], blocks.member_off[b + 1]
        assert blocks.pair[b] == pair and blocks.dir[b] == d and blocks.score[b] == sc, b
        assert (blocks.loc1[a0:a1] == l1).all() and (blocks.loc2[a0:a1] == l2).all(), b
        assert blocks.p_raw[b] == a, (b, blocks.p_raw[b], a)       # fp64 bit-exact (tolerance stated: 0; spec allows 1e-9 rel)


def test_backend_is_cuda(gpu_engine):
    assert gpu_engine.lib.icl_backend() == b"cuda"


def test_golden_clouds_dp_and_blocks(gpu_engine, clouds):
    """Inner seam against the unmodified reference: DP state of both directions, then extraction + p values
    driven by the tie order the reference used."""
    from wgdi_b200.engine import ChainParams
    eng = gpu_engine
    for c in clouds:
        n = len(c["loc1"])
        prm = ChainParams.from_options(c["options"])
        rng = np.random.default_rng(c["case"])
        perm = np.arange(n) if c["case"] % 11 == 5 else rng.permutation(n)      # duplicates keep upload order (H22)
        eng.upload(np.zeros(n, np.int32), c["loc1"][perm], c["loc2"][perm], c["grading"][perm], 1)
        kf, kr = eng.dp(prm)
        assert (kf[::-1] == c["score1"]).all() and (kr[::-1] == c["score2"]).all()      # one pair: reversed segment
        check_dp_against_golden(eng.debug_dp, c)
        blocks = eng.extract(eng.raw_from_order(c["order1"]), eng.raw_from_order(c["order2"]))
        check_blocks_against_golden(blocks, prm.pvalue, c)


def test_all_clouds_as_one_batch(gpu_engine, clouds):
    """The same clouds uploaded together as chromosome pairs of one comparison (shared parameters)."""
    from wgdi_b200.engine import ChainParams
    cs = [c for c in clouds if len(c["loc1"]) > 0]
    pid = np.concatenate([np.full(len(c["loc1"]), i, np.int32) for i, c in enumerate(cs)])
    pts = dict(pair_id=pid, loc1=np.concatenate([c["loc1"] for c in cs]).astype(np.int32),
               loc2=np.concatenate([c["loc2"] for c in cs]).astype(np.int32),
               grading=np.concatenate([c["grading"] for c in cs]).astype(np.int32), n_pairs=len(cs))
    rng = np.random.default_rng(7)
    sh = rng.permutation(len(pid))
    for mg in ((25, 25), (40, 30), (7, 50)):
        prm = ChainParams(mg1=mg[0], mg2=mg[1], grading=(50, 30, 25), coverage_ratio=0.8)
        up = {k: (v[sh] if isinstance(v, np.ndarray) else v) for k, v in pts.items()}
        # equal (loc1, loc2) must keep upload order: sort the shuffle inside duplicates away by using the oracle on
        # the uploaded order itself
        blocks = gpu_engine.chain(up["pair_id"], up["loc1"], up["loc2"], up["grading"], up["n_pairs"], prm)
        _compare_with_oracle(blocks, oracle_blocks(up, prm))


@pytest.mark.parametrize("name", ["tiny", "small", "cross"])
def test_synthetic_comparison_matches_oracle(gpu_engine, name):
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams
    ds = synth.make_dataset(name)
    pts = synth.graded_points(ds)
    c = ds.cfg
    prm = ChainParams(mg1=c.mg[0], mg2=c.mg[1], grading=tuple(c.grading), pvalue=c.pvalue)
    blocks = gpu_engine.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
    exp = oracle_blocks(pts, prm)
    assert len(exp) > 10
    _compare_with_oracle(blocks, exp)
    # DP state as well
    bt, (off, l1, l2, g, o) = oracle_batch(pts, prm)
    for d in (0, 1):
        s, p, dp, u = gpu_engine.debug_dp(d)
        assert (s == bt.dp[d]["score"]).all() and (u == bt.dp[d]["used"]).all()
        # oracle pred is pair-local
        pred_global = np.where(bt.dp[d]["pred"] >= 0, bt.dp[d]["pred"] + np.repeat(off[:-1], np.diff(off)), -1)
        assert (p == pred_global).all() and (dp == bt.dp[d]["depth"]).all()


_E2E_CHILD = r"""
import sys, hashlib, json, io, contextlib, numpy as np
sys.path.insert(0, %r)
fp = hashlib.md5(np.argsort((np.arange(1000) * 7919 %% 5).astype(np.int64), kind="quicksort").tobytes()).hexdigest()
if fp != json.load(open("meta.json"))["argsort_fingerprint"]:
    print("SKIP fingerprint", fp); sys.exit(0)
from wgdi_b200 import _lib, base
from wgdi_b200.run_colliearity import mycollinearity
import os
conf = open("icl.conf").read().replace("expected.collinearity.txt", "got.txt")
if os.environ.get("ICL_TEST_DEVICES"):
    conf += "devices = " + os.environ["ICL_TEST_DEVICES"] + "\n"      # chromosome pairs sharded over several engines
open("icl.conf", "w").write(conf)
with contextlib.redirect_stdout(io.StringIO()):
    try:
        mycollinearity(base.load_conf("icl.conf", "collinearity")).run()
    except SystemExit as e:
        assert e.code == 0
assert _lib.load().icl_backend() == b"cuda"
print("OK" if open("got.txt").read() == open("expected.collinearity.txt").read() else "MISMATCH")
"""


@pytest.mark.parametrize("case", ["tiny", "cross", "tiny_rev", "cross_chrom"])
def test_drop_in_driver_writes_reference_file(case):
    """Outer seam, conf in -> file out, byte-identical to what the unmodified reference wrote."""
    with tempfile.TemporaryDirectory() as d:
        stage_e2e(case, d)
        env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=PORTABLE_ENV)
        r = subprocess.run([sys.executable, "-c", _E2E_CHILD % ROOT], cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        if r.stdout.startswith("SKIP"):
            pytest.skip("numpy scalar-sort fingerprint differs on this machine: " + r.stdout)
        assert r.stdout.strip() == "OK", r.stdout + r.stderr


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except OSError:
        return 0


def test_drop_in_driver_sharded_over_devices():
    """`devices = ...` in the conf: chromosome pairs LPT-sharded over several engines from one process (every visible
    GPU, at least two engines), blocks merged on the host — the file the reference wrote."""
    n = max(2, min(_gpu_count(), 8))
    devices = ",".join(str(i % max(_gpu_count(), 1)) for i in range(n))
    with tempfile.TemporaryDirectory() as d:
        stage_e2e("cross", d)
        env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=PORTABLE_ENV, ICL_TEST_DEVICES=devices)
        r = subprocess.run([sys.executable, "-c", _E2E_CHILD % ROOT], cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        if r.stdout.startswith("SKIP"):
            pytest.skip("numpy scalar-sort fingerprint differs on this machine: " + r.stdout)
        assert r.stdout.strip() == "OK", r.stdout + r.stderr


def test_cli_end_to_end():
    with tempfile.TemporaryDirectory() as d:
        stage_e2e("tiny", d)
        conf = open(os.path.join(d, "icl.conf")).read().replace("expected.collinearity.txt", "got.txt")
        open(os.path.join(d, "icl.conf"), "w").write(conf)
        env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=PORTABLE_ENV, PYTHONPATH=ROOT)
        r = subprocess.run([sys.executable, "-m", "wgdi_b200", "-icl", "icl.conf"], cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "The filtered homologous gene pairs are" in r.stdout and "Running 1 vs 1" in r.stdout
        got = open(os.path.join(d, "got.txt")).read()
        assert got.startswith("# Alignment 1: score=")


def test_inner_seam_class(gpu_engine, clouds):
    """`collinearity(options, points).run()` returns the reference's structure."""
    import pandas as pd
    from oracle import c_oracle, icl_oracle
    from wgdi_b200.collinearity import collinearity
    for c in clouds[10:20]:
        n = len(c["loc1"])
        labels = np.arange(500, 500 + n)[::-1]
        points = pd.DataFrame({"loc1": c["loc1"], "loc2": c["loc2"], "grading": c["grading"]}, index=labels)
        data = collinearity(c["options"], points).run()
        exp = c_oracle.collinearity_run(c["loc1"], c["loc2"], c["grading"], icl_oracle.IclParams(c["options"]))
        assert len(data) == len(exp)
        for (frame, p, score), (b1, b2, ep, esc, _d) in zip(data, exp):
            assert list(frame.columns) == ["loc1", "loc2"]
            assert (frame["loc1"].to_numpy() == b1).all() and (frame["loc2"].to_numpy() == b2).all()
            assert p == ep and score == esc
            # index carries the caller's row labels
            assert (points.loc[frame.index, "loc1"].to_numpy() == b1).all()


def test_edge_cases(gpu_engine):
    from wgdi_b200._lib import IclError
    from wgdi_b200.engine import ChainParams
    eng = gpu_engine
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    z = np.zeros(0, np.int32)
    # empty comparison
    assert len(eng.chain(z, z, z, z, 0, prm)) == 0
    assert len(eng.chain(z, z, z, z, 3, prm)) == 0
    # one point, two points on a diagonal, three points = the smallest block
    one = np.array([0], np.int32)
    assert len(eng.chain(one, one + 5, one + 5, one + 50, 1, prm)) == 0
    tri = eng.chain(np.zeros(3, np.int32), np.array([1, 2, 3], np.int32), np.array([1, 2, 3], np.int32),
                    np.full(3, 50, np.int32), 1, prm)
    assert len(tri) == 1 and tri.m[0] == 3 and tri.score[0] == 50 + 48 + 48 and tri.N[0] == 3
    # pairs without points between pairs with points (ragged input)
    pid = np.array([4, 4, 4, 0, 0, 0], np.int32)
    l = np.array([1, 2, 3, 7, 8, 9], np.int32)
    b = eng.chain(pid, l, l, np.full(6, 50, np.int32), 6, prm)
    assert b.pair.tolist() == [0, 4]
    # anti-diagonal only
    b = eng.chain(np.zeros(4, np.int32), np.array([1, 2, 3, 4], np.int32), np.array([9, 8, 7, 6], np.int32),
                  np.full(4, 50, np.int32), 1, prm)
    assert b.dir.tolist() == [1] and b.loc1[:4].tolist() == [1, 2, 3, 4] and b.loc2[:4].tolist() == [9, 8, 7, 6]
    # error behaviour: codes, not crashes
    with pytest.raises(IclError) as e:
        eng.upload(np.array([2], np.int32), one, one, one, 2)
    assert e.value.code == -1
    with pytest.raises(IclError) as e:
        eng.upload(np.array([0], np.int32), -one - 1, one, one, 1)
    assert e.value.code == -1
    eng.upload(np.zeros(3, np.int32), l[:3], l[:3], np.full(3, 50, np.int32), 1)
    with pytest.raises(IclError) as e:
        eng.extract(np.zeros(3, np.int32), np.zeros(3, np.int32))
    assert e.value.code == -3                                   # extract before dp
    with pytest.raises(IclError) as e:
        eng.dp(ChainParams(mg1=300, mg2=25))
    assert e.value.code == -4
    eng.dp(prm)
    with pytest.raises(IclError) as e:
        eng.extract(np.array([0, 1, 7], np.int32), np.array([0, 1, 2], np.int32))
    assert e.value.code == -1                                   # argsort index outside its pair
    with pytest.raises(IclError) as e:
        eng.extract(np.array([0, 1, 1], np.int32), np.array([0, 1, 2], np.int32))
    assert e.value.code == -1                                   # not a permutation


def test_documented_limits_are_errors(gpu_engine):
    from wgdi_b200._lib import IclError
    from wgdi_b200.engine import ChainParams, IclEngine
    one = np.array([0], np.int32)
    with pytest.raises(IclError) as e:
        gpu_engine.upload(one, one + 1, one + 1, one + 300, 1)          # grading > 255
    assert e.value.code == -4
    with pytest.raises(IclError) as e:
        gpu_engine.upload(one, one + (1 << 26), one + 1, one + 50, 1)   # coordinate >= 2^26
    assert e.value.code == -4


def test_window_bounds_are_exclusive(gpu_engine):
    """Offsets 1..mg-1 chain, offset mg does not (collinearity.py:76-79), in both directions."""
    from wgdi_b200.engine import ChainParams
    for mg1, mg2 in ((5, 9), (9, 5)):
        prm = ChainParams(mg1=mg1, mg2=mg2, grading=(50, 30, 25))
        for dr, dc, linked in ((mg1 - 1, mg2 - 1, True), (mg1, 1, False), (1, mg2, False)):
            l1 = np.array([10, 10 + dr], np.int32)
            l2 = np.array([10, 10 + dc], np.int32)
            gpu_engine.upload(np.zeros(2, np.int32), l1, l2, np.full(2, 50, np.int32), 1)
            kf, _ = gpu_engine.dp(prm)
            assert (kf[0] == 100 - dr - dc) == linked and (linked or kf[0] == 50)       # keys are reversed per segment


def test_forked_chains(gpu_engine):
    """Long paths sharing a root-side prefix: exercises the removed-ancestor memo of the extraction kernel."""
    from conftest import forked_chain_points
    from wgdi_b200.engine import ChainParams
    for prefix, arm in ((60, 200), (400, 900)):
        pts = forked_chain_points(prefix, arm)
        for cov in (0.8, 0.5):
            prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25), coverage_ratio=cov)
            blocks = gpu_engine.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], 1, prm)
            _compare_with_oracle(blocks, oracle_blocks(pts, prm))


def test_sharded_equals_unsharded(gpu_engine):
    """Chromosome pairs split over 2 and 3 ranks by LPT (ranks emulated one after another on this GPU)."""
    from wgdi_b200 import shard, synth
    from wgdi_b200.engine import ChainParams
    pts = synth.graded_points(synth.make_dataset("small"))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    whole = gpu_engine.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
    for world in (2, 3):
        parts = []
        for r in range(world):
            sp = shard.shard_points(pts, r, world)
            b = gpu_engine.chain(sp["pair_id"], sp["loc1"], sp["loc2"], sp["grading"], sp["n_pairs"], prm)
            b.pair = sp["global_pair"][b.pair]
            b.member_idx = sp["rows"][b.member_idx].astype(np.int32)
            parts.append(b)
        merged = shard.merge_blocks(parts)
        for f in ("pair", "dir", "score", "m", "L", "N", "N1", "L1", "L2", "p_raw", "member_off", "loc1", "loc2", "member_idx"):
            assert (getattr(merged, f) == getattr(whole, f)).all(), f


@pytest.mark.parametrize("name,scale,mg", [("C1", 1.0, (25, 25)), ("C3", 1.0, (25, 25)), ("C4", 1.0, (25, 25)),
                                           ("C3", 0.3, (40, 40)), ("C4", 0.25, (50, 50))])
def test_other_baseline_configs_match_oracle(gpu_engine, name, scale, mg):
    """BASELINE.json configs 1 (25 pairs of ~40k points), 3 (cross-species, ~4.7 M hits, 144 pairs) and 4 (wheat-like,
    ~9.9 M hits, 441 pairs) at FULL size with their own parameters, plus 3 and 4 at a bounded scale with the other two
    `mg` settings of config 5's sweep: block lists identical to the C oracle, block by block."""
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams
    ds = synth.make_dataset(name, scale=scale)
    pts = synth.graded_points(ds)
    prm = ChainParams(mg1=mg[0], mg2=mg[1], grading=tuple(ds.cfg.grading))
    blocks = gpu_engine.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
    _compare_with_oracle(blocks, oracle_blocks(pts, prm))
    assert gpu_engine.counters()["chain_kernel"] == "ring"


def test_streamed_rank_equals_two_step_rank(gpu_engine):
    """icl_dp_begin (items export their scores and raise a flag as they finish, the sort pool follows) gives the
    same score arrays and the same numpy argsort as icl_dp followed by the rank, including empty chromosome pairs."""
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams
    pts = synth.graded_points(synth.make_dataset("small"))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    pair = pts["pair_id"].astype(np.int32) * 2                      # every second pair id is empty
    gpu_engine.upload(pair, pts["loc1"], pts["loc2"], pts["grading"], 2 * pts["n_pairs"])
    kf, kr = gpu_engine.dp(prm)
    kf, kr = kf.copy(), kr.copy()
    af, ar = [x.copy() for x in gpu_engine.rank(kf, kr)]
    for _ in range(3):
        sf, sr = gpu_engine.dp_rank_streamed(prm)
        keys = gpu_engine._buf("keys", 2 * gpu_engine.n, np.int32)
        assert (keys[:gpu_engine.n] == kf).all() and (keys[gpu_engine.n:] == kr).all()
        assert (sf == af).all() and (sr == ar).all()
    assert gpu_engine.counters()["chain_kernel"] == "ring"
    b1 = gpu_engine.extract(sf, sr)
    _compare_with_oracle(b1, [(2 * p, d, l1, l2, a, sc) for p, d, l1, l2, a, sc in oracle_blocks(pts, prm)])


def test_batch_runner_equals_one_engine(gpu_engine):
    """Config 5 shape: several comparisons x a grading x mg sweep through two engines on one GPU; every job's blocks
    equal those of a lone engine fed points graded from scratch with that job's `grading`, and the oracle's on the same
    independently graded input (a grading sweep changes every point's value: run_colliearity.py:30-33,64-86)."""
    from wgdi_b200 import synth
    from wgdi_b200.batch import BatchRunner, sweep_jobs
    dss = [synth.make_dataset(name) for name in ("tiny", "cross", "small")]
    comps = [synth.graded_points(ds) for ds in dss]
    gradings = [(50, 30, 25), (50, 40, 25)]
    jobs, keys = [], []
    for ci, pts in enumerate(comps):
        j, k = sweep_jobs(pts, gradings, [(25, 25), (40, 40), (50, 50)], [0.05, 0.2, 1])
        assert len(j) == 6 and len(k[0][2]) == 3
        jobs += j
        keys += [(ci,) + kk for kk in k]
    runner = BatchRunner(0, depth=2, sort_threads=4)
    try:
        got = runner.run(jobs, min_members=1)
    finally:
        runner.close()
    assert len(got) == len(jobs) == 18
    for (_, prm), (ci, g, mg, _), b in zip(jobs, keys, got):
        own = synth.graded_points(dss[ci], grading=g)
        exp = gpu_engine.chain(own["pair_id"], own["loc1"], own["loc2"], own["grading"], own["n_pairs"], prm)
        assert len(b) == len(exp)
        for name in ("pair", "dir", "score", "m", "N", "N1", "p_raw", "member_off", "loc1", "loc2", "member_idx"):
            assert (getattr(b, name) == getattr(exp, name)).all(), name
    for j in (0, 3, 8, 15):
        ci, g = keys[j][0], keys[j][1]
        _compare_with_oracle(got[j], oracle_blocks(synth.graded_points(dss[ci], grading=g), jobs[j][1]))
    # 50,40,25 against 50,30,25 at the same mg: different blue values, different result
    assert len(got[0]) != len(got[3]) or (got[0].score != got[3].score).any()


def test_config5_comparison_full_sweep_matches_oracle(gpu_engine):
    """BASELINE.json config 5: one of the 28 cross-species comparisons at full size through the batch runner with all
    six (grading, mg) device jobs of the sweep — the three pvalue variants of a job share its blocks — every job
    identical to the oracle run on points graded FROM SCRATCH with that job's grading (run_colliearity.py:64-86), and the
    two gradings give different results at every mg."""
    from wgdi_b200 import synth
    from wgdi_b200.batch import BatchRunner, sweep_jobs
    ds = synth.make_dataset(synth.config5_comparison(9))
    pts = synth.graded_points(ds)
    assert len(pts["loc1"]) > 1_000_000
    jobs, keys = sweep_jobs(pts, synth.C5_GRADINGS, synth.C5_MGS, synth.C5_PVALUES)
    assert len(jobs) == 6
    runner = BatchRunner(0, depth=2)
    try:
        got = runner.run(jobs, min_members=1)
    finally:
        runner.close()
    own = {g: synth.graded_points(ds, grading=g) for g in synth.C5_GRADINGS}
    assert (own[(50, 40, 25)]["grading"] != own[(50, 30, 25)]["grading"]).any()
    for (_, prm), (g, mg, pvs), b in zip(jobs, keys, got):
        exp = oracle_blocks(own[g], prm)
        _compare_with_oracle(b, exp)
        p = b.p_value()
        kept = [int((~(p > pv) & (b.m >= 5)).sum()) for pv in pvs]            # what the writer keeps per pvalue variant
        assert 0 < kept[0] <= kept[1] <= kept[2] <= int((b.m >= 5).sum())     # (the unrounded `a` may exceed 1)
    for j in range(3):                                                          # same mg, the two gradings
        assert len(got[j]) != len(got[3 + j]) or (got[j].score != got[3 + j].score).any()


def _random_comparison(rng, n_pairs, n_rows, n_cols, per_row, diag_frac, dup_frac):
    """Random chromosome pairs with a mix of diagonal runs (both orientations, with gaps and side hits) and noise;
    several hits per row and duplicate coordinates allowed."""
    pid, l1, l2 = [], [], []
    for p in range(n_pairs):
        n = int(n_rows * per_row)
        rows = rng.integers(1, n_rows + 1, n)
        cols = rng.integers(1, n_cols + 1, n)
        nd = int(n * diag_frac)
        k = 0
        while k < nd:
            r, c = int(rng.integers(1, n_rows)), int(rng.integers(1, n_cols))
            sign = 1 if rng.random() < 0.5 else -1
            for _ in range(int(rng.integers(3, 60))):
                if k >= nd or not (1 <= r <= n_rows and 1 <= c <= n_cols):
                    break
                rows[k], cols[k] = r, c
                k += 1
                r += int(rng.integers(1, 4))
                c += sign * int(rng.integers(1, 4))
        nd2 = int(n * dup_frac)
        if nd2:
            src = rng.integers(0, n, nd2)
            rows, cols = np.r_[rows, rows[src]], np.r_[cols, cols[src]]
        pid.append(np.full(len(rows), p)); l1.append(rows); l2.append(cols)
    pid, l1, l2 = (np.concatenate(x).astype(np.int32) for x in (pid, l1, l2))
    g = rng.choice(np.array([50, 30, 25], np.int32), len(pid), p=[0.25, 0.35, 0.4])
    o = rng.permutation(len(pid))
    return dict(pair_id=pid[o], loc1=l1[o], loc2=l2[o], grading=g[o], n_pairs=n_pairs)


@pytest.mark.parametrize("seed", range(9))
def test_random_comparisons_match_oracle(gpu_engine, seed):
    """Randomised parity: dense and sparse rows, mg from 2 to 255 (window spans from a few points to beyond the
    shared-memory ring, i.e. every ring size and the global-memory fallback), odd coverage ratios, duplicates."""
    from wgdi_b200.engine import ChainParams
    rng = np.random.default_rng(1000 + seed)
    shapes = [  # (pairs, rows, cols, hits per row, diagonal share, duplicate share, mg1, mg2)
        (3, 400, 400, 6.0, 0.5, 0.02, 25, 25), (2, 150, 900, 40.0, 0.3, 0.0, 60, 120), (4, 2000, 2000, 1.5, 0.7, 0.01, 2, 2),
        (1, 120, 120, 40.0, 0.2, 0.0, 255, 255), (2, 1200, 300, 3.0, 0.6, 0.05, 5, 200), (3, 800, 800, 8.0, 0.4, 0.0, 120, 9),
        (1, 60, 60, 30.0, 0.5, 0.1, 40, 40), (5, 500, 700, 2.5, 0.8, 0.0, 33, 64), (1, 140, 140, 70.0, 0.1, 0.0, 255, 255)]
    npairs, nr, nc, per, diag, dup, mg1, mg2 = shapes[seed]
    pts = _random_comparison(rng, npairs, nr, nc, per, diag, dup)
    prm = ChainParams(mg1=mg1, mg2=mg2, grading=(50, 30, 25), coverage_ratio=float(rng.choice([0.8, 0.5, 0.95])))
    blocks = gpu_engine.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
    exp = oracle_blocks(pts, prm)
    assert len(exp) > 0
    _compare_with_oracle(blocks, exp)
    assert gpu_engine.counters()["chain_kernel"] == ("generic" if seed == 8 else "ring")     # seed 8: window beyond the ring
    # the DP state itself, point by point (score, predecessor, path length, used) against the C oracle
    from oracle import c_oracle
    off, perm = gpu_engine.layout()
    for d in (0, 1):
        s, pr, dp, us = gpu_engine.debug_dp(d)
        for p in range(pts["n_pairs"]):
            a, b = off[p], off[p + 1]
            if a == b:
                continue
            i = perm[a:b]
            es, ep, eu, ed = c_oracle.chain(pts["loc1"][i], pts["loc2"][i], pts["grading"][i], mg1, mg2, 1 if d == 0 else -1)
            assert (s[a:b] == es).all() and (us[a:b].astype(bool) == eu).all() and (dp[a:b] == ed).all(), (seed, d, p)
            assert (np.where(pr[a:b] >= 0, pr[a:b] - a, -1) == ep).all(), (seed, d, p)


@pytest.mark.parametrize("name,scale", [("small", 1.0), ("C2", 0.1)])
def test_device_rank_mode_matches_stable_oracle(gpu_engine, name, scale):
    """Non-reference option (icl_rank_device): extraction order from a stable device sort; bit-exact against the oracle
    run with the same stable order."""
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams
    pts = synth.graded_points(synth.make_dataset(name, scale=scale))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    gpu_engine.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"])
    got = gpu_engine.run_resident_device_rank(prm)
    _compare_with_oracle(got, oracle_blocks_stable(pts, prm))


_NPSCALAR_CHILD = r"""
import sys, ctypes as C, numpy as np
sys.path.insert(0, %r)
from wgdi_b200 import synth
from wgdi_b200._lib import icl_params
from wgdi_b200.engine import ChainParams, IclEngine
KIND = sys.argv[1]
MODE = {"numpy_scalar": 1, "numpy_avx512": 2, "numpy_avx2": 3}[KIND]
def check(pts, prm):
    eng = IclEngine(0)
    try:
        eng.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"])
        kf, kr = eng.dp(prm)                                  # the arrays the host would sort
        n = eng.n
        keys = np.concatenate([kf, kr])
        off = np.concatenate([eng.pair_off, eng.pair_off[1:] + n])
        eng._check(eng.lib.icl_rank_device(eng.ctx, MODE))    # numpy's argsort of this flavour, on the device
        got = np.zeros(2 * n, np.int32)
        eng._check(eng.lib.icl_debug_order(eng.ctx, got.ctypes.data))
        bad = 0
        for a, b in zip(off[:-1], off[1:]):
            exp = keys[a:b].astype(np.int64).argsort(kind="quicksort")      # numpy in THIS process: SIMD sorts disabled
            bad += not (got[a:b] == exp).all()
        b1 = eng.run_resident_device_rank(prm, mode=KIND)
        assert eng.host_rank_fallbacks == 0
        eng.rank_mode = "host"
        b2 = eng.run_resident(prm)                            # host rank with the same numpy
        same = len(b1) == len(b2) and all((getattr(b1, f) == getattr(b2, f)).all() for f in
                                          ("pair", "dir", "score", "m", "N", "N1", "p_raw", "member_off", "loc1", "loc2", "member_idx"))
        return len(off) - 1, bad, same, eng.numpy_sort_kind() == KIND
    finally:
        eng.close()
rng = np.random.default_rng(3)
out = []
out.append(check(synth.graded_points(synth.make_dataset("small")), ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))))
out.append(check(synth.graded_points(synth.make_dataset("C2", scale=0.15)), ChainParams(mg1=40, mg2=40, grading=(50, 40, 25))))
# adversarial segments: all ties, sorted, reversed, organ pipe, few distinct values, one huge pair
n = 40000
l1 = np.repeat(np.arange(1, n // 4 + 1), 4).astype(np.int32); l2 = rng.integers(1, 3000, n).astype(np.int32)
out.append(check(dict(pair_id=np.zeros(n, np.int32), loc1=l1, loc2=l2, grading=rng.choice(np.array([50, 30, 25], np.int32), n), n_pairs=1),
                 ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))))
iso = np.arange(1, 901, dtype=np.int32) * 1000                  # isolated points: constant / sorted / reversed key segments
out.append(check(dict(pair_id=np.repeat(np.arange(3, dtype=np.int32), 300), loc1=iso, loc2=rng.permutation(iso).astype(np.int32),
                      grading=np.concatenate([np.full(300, 50), np.sort(rng.choice([25, 30, 50], 300)),
                                              np.sort(rng.choice([25, 30, 50], 300))[::-1]]).astype(np.int32), n_pairs=3),
                 ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))))
print("RESULT", out)
assert all(bad == 0 and same and scalar for _, bad, same, scalar in out), out
print("OK")
"""


def test_device_numpy_scalar_argsort_is_numpys():
    """icl_rank_device(ICL_RANK_NUMPY_SCALAR) against numpy itself in a process whose numpy runs its scalar introsort
    (NPY_DISABLE_CPU_FEATURES): every segment's argsort identical, and the blocks identical to the host-ranked path."""
    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=PORTABLE_ENV)
    r = subprocess.run([sys.executable, "-c", _NPSCALAR_CHILD % ROOT, "numpy_scalar"], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


AVX2_ENV = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX512_KNL AVX512_KNM"


def test_device_numpy_avx2_argsort_is_numpys():
    """icl_rank_device(ICL_RANK_NUMPY_AVX2) against numpy itself in a process whose numpy dispatches int64 argsort to its
    AVX2 kernels (every AVX512* feature disabled): every segment's argsort identical, the engine's start-up probe picks
    that flavour, and the blocks are identical to the host-ranked path."""
    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=AVX2_ENV)
    probe = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import xss_emulation as x; "
                            "print('AVX2' if x.is_avx2_numpy() else 'OTHER')" % os.path.join(ROOT, "tools")],
                           env=env, capture_output=True, text=True)
    if "AVX2" not in probe.stdout:
        pytest.skip("numpy on this machine has no AVX2 argsort to compare with")
    r = subprocess.run([sys.executable, "-c", _NPSCALAR_CHILD % ROOT, "numpy_avx2"], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_device_numpy_avx512_argsort_is_numpys(gpu_engine):
    """icl_rank_device(ICL_RANK_NUMPY_AVX512) against numpy in this process, where numpy sorts int64 with its AVX-512
    kernels: every segment's argsort identical (all sizes 1..600, real score arrays, constant / sorted segments, segments
    of every size class of the kernel's launches), and the blocks identical to the host-ranked path."""
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams, IclEngine
    if gpu_engine.numpy_sort_kind() != "numpy_avx512":
        pytest.skip("numpy in this process does not use its AVX-512 argsort")
    rng = np.random.default_rng(8)
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    sizes = list(range(1, 601, 5))
    cases = [dict(pair_id=np.repeat(np.arange(len(sizes), dtype=np.int32), sizes),
                  loc1=np.concatenate([np.arange(1, m + 1) for m in sizes]).astype(np.int32),
                  loc2=rng.integers(1, 40, sum(sizes)).astype(np.int32),
                  grading=rng.choice(np.array([50, 30, 25], np.int32), sum(sizes)), n_pairs=len(sizes)),
             synth.graded_points(synth.make_dataset("small")),
             synth.graded_points(synth.make_dataset("C2", scale=0.15))]
    # one launch per size class of the rank kernel: <= 16 000 keys (above), <= 34 000 (keys + 16-bit indices in up to 204 KB
    # of shared memory), <= 65 535 (indices in shared memory, keys through L1), longer (32-bit indices in global memory);
    # the last case mixes three classes in one comparison
    for n in (40000, 20000, 70000):
        cases.append(dict(pair_id=np.zeros(n, np.int32), loc1=np.repeat(np.arange(1, n // 4 + 1), 4).astype(np.int32),
                          loc2=rng.integers(1, 3000, n).astype(np.int32), grading=rng.choice(np.array([50, 30, 25], np.int32), n), n_pairs=1))
    sizes3 = [70000, 36000, 17000, 900, 300]
    cases.append(dict(pair_id=np.repeat(np.arange(5, dtype=np.int32), sizes3),
                      loc1=np.concatenate([np.repeat(np.arange(1, m // 4 + 1), 4) for m in sizes3]).astype(np.int32),
                      loc2=rng.integers(1, 3000, sum(sizes3)).astype(np.int32),
                      grading=rng.choice(np.array([50, 30, 25], np.int32), sum(sizes3)), n_pairs=5))
    # isolated points: every score equals the grading -> constant and sorted key segments (numpy's SIMD argsort leaves
    # an already sorted array untouched; a partition pass whose pivot is the smallest key moves everything)
    iso = np.arange(1, 901, dtype=np.int32) * 1000
    cases.append(dict(pair_id=np.repeat(np.arange(3, dtype=np.int32), 300), loc1=iso, loc2=rng.permutation(iso).astype(np.int32),
                      grading=np.concatenate([np.full(300, 50), np.sort(rng.choice([25, 30, 50], 300)),
                                              np.sort(rng.choice([25, 30, 50], 300))[::-1]]).astype(np.int32), n_pairs=3))
    eng = IclEngine(0)
    try:
        for pts in cases:
            eng.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"])
            kf, kr = eng.dp(prm)
            m = eng.n
            keys = np.concatenate([kf, kr])
            off = np.concatenate([eng.pair_off, eng.pair_off[1:] + m])
            eng._check(eng.lib.icl_rank_device(eng.ctx, 2))
            got = np.zeros(2 * m, np.int32)
            eng._check(eng.lib.icl_debug_order(eng.ctx, got.ctypes.data))
            for a, b in zip(off[:-1], off[1:]):
                assert (got[a:b] == keys[a:b].astype(np.int64).argsort(kind="quicksort")).all(), (a, b)
            b1 = eng.run_resident_device_rank(prm, mode="numpy_avx512")
            assert eng.host_rank_fallbacks == 0
            eng.rank_mode = "host"
            b2 = eng.run_resident(prm)
            eng.rank_mode = "auto"
            assert len(b1) == len(b2)
            for f in ("pair", "dir", "score", "m", "N", "N1", "p_raw", "member_off", "loc1", "loc2", "member_idx"):
                assert (getattr(b1, f) == getattr(b2, f)).all(), f
    finally:
        eng.close()


def test_device_rank_is_deterministic(gpu_engine):
    """The rank kernel's warps hand sub-ranges to each other through a shared stack; whatever order they take them in,
    the permutation is a function of the keys alone: 25 repetitions (and 1, 2, 4 warps per segment in child processes)
    give the same order as numpy every time."""
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams, IclEngine
    kind = gpu_engine.numpy_sort_kind()
    if kind is None:
        pytest.skip("numpy in this process is none of the reproduced flavours")
    mode = {"numpy_scalar": 1, "numpy_avx512": 2, "numpy_avx2": 3}[kind]
    pts = synth.graded_points(synth.make_dataset("C2", scale=0.15))
    prm = ChainParams(mg1=40, mg2=40, grading=(50, 40, 25))
    eng = IclEngine(0)
    try:
        eng.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"])
        kf, kr = eng.dp(prm)
        n = eng.n
        keys = np.concatenate([kf, kr])
        off = np.concatenate([eng.pair_off, eng.pair_off[1:] + n])
        exp = np.concatenate([keys[a:b].astype(np.int64).argsort(kind="quicksort") for a, b in zip(off[:-1], off[1:])])
        got = np.zeros(2 * n, np.int32)
        for rep in range(25):
            eng._check(eng.lib.icl_rank_device(eng.ctx, mode))
            eng._check(eng.lib.icl_debug_order(eng.ctx, got.ctypes.data))
            assert (got == exp).all(), rep
    finally:
        eng.close()
    if kind != "numpy_scalar":
        child = ("import sys; sys.path.insert(0, %r)\n"
                 "import numpy as np\nfrom wgdi_b200 import synth\nfrom wgdi_b200.engine import ChainParams, IclEngine\n"
                 "pts = synth.graded_points(synth.make_dataset('C2', scale=0.15))\n"
                 "prm = ChainParams(mg1=40, mg2=40, grading=(50, 40, 25))\n"
                 "eng = IclEngine(0); eng.upload(pts['pair_id'], pts['loc1'], pts['loc2'], pts['grading'], pts['n_pairs'])\n"
                 "b1 = eng.run_resident(prm); assert eng.numpy_sort_kind() == %r and eng.host_rank_fallbacks == 0\n"
                 "eng.rank_mode = 'host'; b2 = eng.run_resident(prm)\n"
                 "assert len(b1) == len(b2) and all((getattr(b1, f) == getattr(b2, f)).all() for f in ('pair', 'dir', 'score', 'p_raw', 'loc1', 'loc2'))\n"
                 "print('OK')\n" % (ROOT, kind))
        for warps in (1, 2, 4):
            r = subprocess.run([sys.executable, "-c", child], env=dict(os.environ, ICL_B200_XS_WARPS=str(warps)),
                               capture_output=True, text=True)
            assert r.returncode == 0 and "OK" in r.stdout, (warps, r.stdout[-1000:] + r.stderr[-2000:])


def test_alternative_dp_kernels_agree(clouds):
    """The alternative DP kernels (ICL_B200_CHAIN=generic / single) give the same blocks as the default ring kernel."""
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams, IclEngine
    pts = synth.graded_points(synth.make_dataset("small"))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    exp = oracle_blocks(pts, prm)
    for mode in ("generic", "single"):
        os.environ["ICL_B200_CHAIN"] = mode
        try:
            eng = IclEngine(0)
            blocks = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
            assert eng.counters()["chain_kernel"] == mode
            eng.close()
        finally:
            del os.environ["ICL_B200_CHAIN"]
        _compare_with_oracle(blocks, exp)


@pytest.mark.parametrize("mode", ["seq", "inc"])
def test_both_pvalue_paths_agree_with_oracle(mode):
    """The sequential warp-per-pair `times` kernel and the order-free incidence path, both forced on one input."""
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams, IclEngine
    pts = synth.graded_points(synth.make_dataset("small"))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    os.environ["ICL_B200_PVALUE"] = mode
    try:
        eng = IclEngine(0)
        blocks = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
        eng.close()
    finally:
        del os.environ["ICL_B200_PVALUE"]
    _compare_with_oracle(blocks, oracle_blocks(pts, prm))


def test_full_size_C2_matches_oracle_and_properties(gpu_engine):
    """BASELINE.json config 2 (Vitis-sized self comparison, ~2M graded hits) at full size: identical to the C oracle,
    plus size-independent properties."""
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams
    ds = synth.make_dataset("C2")
    pts = synth.graded_points(ds)
    assert pts["n_pairs"] == 361 and len(pts["loc1"]) > 1_500_000
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    blocks = gpu_engine.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
    again = gpu_engine.run_resident(prm)                          # idempotent on resident inputs
    for f in ("pair", "dir", "score", "m", "p_raw", "member_off", "loc1", "loc2"):
        assert (getattr(again, f) == getattr(blocks, f)).all()
    _compare_with_oracle(blocks, oracle_blocks(pts, prm))
    # properties: members ascend on loc1; forward blocks ascend / anti-diagonal descend on loc2; every point is in at
    # most one block per direction; m <= L; N >= m; blocks ordered by (pair, dir)
    seg = np.repeat(np.arange(len(blocks)), blocks.m)
    d1 = np.diff(blocks.loc1.astype(np.int64)); d2 = np.diff(blocks.loc2.astype(np.int64)); same = np.diff(seg) == 0
    assert (d1[same] > 0).all()
    assert (d2[same & (blocks.dir[seg][1:] == 0)] > 0).all() and (d2[same & (blocks.dir[seg][1:] == 1)] < 0).all()
    key = blocks.dir[seg].astype(np.int64) * len(pts["loc1"]) + blocks.member_idx
    assert len(np.unique(key)) == len(key)
    assert (blocks.m <= blocks.L).all() and (blocks.N >= blocks.m).all() and (blocks.m >= 3).all()
    order_key = blocks.pair.astype(np.int64) * 2 + blocks.dir
    assert (np.diff(order_key) >= 0).all()


def test_rank_probe_covers_every_launch_class(gpu_engine):
    """The start-up probe that lets the device rank sorts one tie-heavy segment per launch size class of the rank kernels
    (<= 16 000 keys, <= 34 000, <= 65 535, longer) with the device and with the running numpy — and it discriminates:
    the other SIMD flavour fails it."""
    from wgdi_b200.engine import IclEngine
    kind = gpu_engine.numpy_sort_kind()
    if kind is None:
        pytest.skip("numpy in this process is none of the reproduced flavours")
    mode = {"numpy_scalar": 1, "numpy_avx512": 2, "numpy_avx2": 3}[kind]
    sizes = IclEngine.PROBE_CLASS_SIZES
    assert max(sizes) > 65535 and any(34000 < m <= 65535 for m in sizes) and any(16000 < m <= 34000 for m in sizes)
    assert gpu_engine._probe_size_classes(mode)
    other = {1: 2, 2: 3, 3: 2}[mode]
    assert not gpu_engine._probe_size_classes(other)
    # the probe entry point on segments of every size 0..300 and on empty input
    rng = np.random.default_rng(5)
    segs = [IclEngine._score_like_keys(rng, m, m % 3) for m in range(0, 300)]
    off = np.r_[0, np.cumsum([len(x) for x in segs])].astype(np.int64)
    keys = np.concatenate(segs).astype(np.int32)
    got = np.zeros(len(keys), np.int32)
    gpu_engine._check(gpu_engine.lib.icl_debug_rank_keys(gpu_engine.ctx, mode, len(segs), off.ctypes.data, keys.ctypes.data, got.ctypes.data))
    for a, b, x in zip(off[:-1], off[1:], segs):
        assert (got[a:b] == x.astype(np.int64).argsort(kind="quicksort")).all(), b - a
    assert gpu_engine.lib.icl_debug_rank_keys(gpu_engine.ctx, mode, 0, off.ctypes.data, None, None) == 0
    assert gpu_engine.lib.icl_debug_rank_keys(gpu_engine.ctx, 0, 1, off.ctypes.data, keys.ctypes.data, got.ctypes.data) == -1


def test_rank_verify_switch_and_forced_hand_back(gpu_engine, monkeypatch):
    """ICL_B200_RANK_VERIFY re-sorts a sample of the device-ranked segments with numpy every step; a rank that is not
    numpy's (forced here: the other SIMD flavour) is caught, the comparison goes to the host rank and the blocks are the
    host-ranked ones.  ICL_B200_XS_FORCE_GIVEUP drives the rank kernel's give-up path (s_giveup -> host_flag ->
    ICL_ERR_STATE) and the engine's hand-back to numpy."""
    from wgdi_b200 import engine, synth
    from wgdi_b200.engine import ChainParams, IclEngine
    kind = gpu_engine.numpy_sort_kind()
    if kind is None:
        pytest.skip("numpy in this process is none of the reproduced flavours")
    pts = synth.graded_points(synth.make_dataset("small"))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    gpu_engine.rank_mode = "host"
    try:
        ref = gpu_engine.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
    finally:
        gpu_engine.rank_mode = "auto"
    fields = ("pair", "dir", "score", "m", "N", "N1", "p_raw", "member_off", "loc1", "loc2", "member_idx")
    monkeypatch.setenv("ICL_B200_RANK_VERIFY", "1")
    eng = IclEngine(0)
    try:
        assert eng.rank_verify == 32
        got = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
        assert eng.rank_verified_segments >= 8 and eng.rank_verify_failures == 0 and eng.host_rank_fallbacks == 0
        assert all((getattr(got, f) == getattr(ref, f)).all() for f in fields)
        if kind != "numpy_scalar":
            eng.rank_mode = "numpy_avx2" if kind == "numpy_avx512" else "numpy_avx512"      # not this numpy's order
            got = eng.run_resident(prm)
            assert eng.rank_verify_failures == 1 and engine._NP_KIND[0] is None
            assert all((getattr(got, f) == getattr(ref, f)).all() for f in fields)          # ranked by numpy after all
            eng.rank_mode = "auto"
            got = eng.run_resident(prm)                                                     # the process stays on the host rank
            assert all((getattr(got, f) == getattr(ref, f)).all() for f in fields)
    finally:
        eng.close()
        engine._NP_KIND.clear()
    monkeypatch.delenv("ICL_B200_RANK_VERIFY")
    if kind != "numpy_scalar":
        monkeypatch.setenv("ICL_B200_XS_FORCE_GIVEUP", "1")
        eng = IclEngine(0)
        try:
            eng.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"])
            got = eng.run_resident_device_rank(prm, mode=kind)
            assert eng.host_rank_fallbacks == 1
            assert all((getattr(got, f) == getattr(ref, f)).all() for f in fields)
        finally:
            eng.close()


def test_typed_upload_equals_int32_upload(gpu_engine):
    """icl_upload_typed: the point arrays in uint8 / uint16 (what params.compact_points hands over: 7 bytes per point on
    the wire instead of 16) give exactly the blocks of the int32 upload; mixed widths, and values that do not fit stay int32."""
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams
    from wgdi_b200.params import compact_points
    pts = synth.graded_points(synth.make_dataset("small"))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    ref = gpu_engine.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
    fields = ("pair", "dir", "score", "m", "N", "N1", "p_raw", "member_off", "loc1", "loc2", "member_idx")
    cp = compact_points(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"])
    assert [a.dtype for a in cp] == [np.uint8, np.uint16, np.uint16, np.uint8]
    got = gpu_engine.chain(*cp, pts["n_pairs"], prm)
    assert gpu_engine.upload_bytes == 6 * len(pts["loc1"])
    assert all((getattr(got, f) == getattr(ref, f)).all() for f in fields)
    mixed = (pts["pair_id"].astype(np.uint16), pts["loc1"].astype(np.int32), pts["loc2"].astype(np.uint16), pts["grading"].astype(np.uint8))
    got = gpu_engine.chain(*mixed, pts["n_pairs"], prm)
    assert all((getattr(got, f) == getattr(ref, f)).all() for f in fields)
    big = compact_points(np.array([0, 70000]), np.array([1, 2]), np.array([3, 4]), np.array([50, 300]))
    assert [a.dtype for a in big] == [np.int32, np.uint8, np.uint8, np.uint16]
    from wgdi_b200._lib import IclError
    with pytest.raises(IclError) as e:
        gpu_engine.upload(np.zeros(2, np.uint8), big[1], big[2], big[3], 1)          # grading 300 > 255
    assert e.value.code == -4
    assert gpu_engine.lib.icl_upload_typed(gpu_engine.ctx, 2, 1, big[1].ctypes.data, 3, big[1].ctypes.data, 1, big[2].ctypes.data, 1,
                                           big[1].ctypes.data, 1) == -1              # element size 3


@pytest.mark.gpu
def test_min_members_selection_is_a_filter(gpu_engine):
    """Full C2 size: the output for min_members = 5 (what the file writer takes) is the full result filtered."""
    from test_host_logic import _assert_selection_is_a_filter
    from bench import make_points, params_for
    pts = make_points("C2")
    _assert_selection_is_a_filter(gpu_engine, pts, params_for(pts["cfg"]), 5)


def _random_pairs_points(n_pairs, n_points, span, seed, offset=0):
    """Short diagonals and noise spread over many chromosome pairs (some pairs stay empty)."""
    rng = np.random.default_rng(seed)
    pair = rng.integers(0, n_pairs, n_points // 8, dtype=np.int64)
    start1 = rng.integers(1, span, len(pair))
    start2 = rng.integers(1, span, len(pair))
    sign = rng.choice([-1, 1], len(pair))
    k = np.arange(6)
    p = np.repeat(pair, 6)
    l1 = (start1[:, None] + k[None, :]).ravel()
    l2 = (start2[:, None] + sign[:, None] * k[None, :]).ravel()
    nz = n_points - len(p)
    p = np.r_[p, rng.integers(0, n_pairs, nz)]
    l1 = np.r_[l1, rng.integers(1, span, nz)]
    l2 = np.r_[l2, rng.integers(1, span, nz)]
    ok = l2 > 0
    g = rng.choice([50, 30, 25], len(p))[ok]
    return dict(pair_id=p[ok].astype(np.int32), loc1=(l1[ok] + offset).astype(np.int32), loc2=(l2[ok] + offset).astype(np.int32),
                grading=g.astype(np.int32), n_pairs=n_pairs)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["many_pairs", "wide_keys"])
def test_upload_layout_paths(gpu_engine, case):
    """The layout planned at upload, on its two other paths: more pairs than the shared-memory histogram of the validation
    pass holds (2048: per-pair counters by global atomics) and coordinates that need 64-bit sort keys (loc up to 2^26 - 1);
    blocks against the oracle, layout against a numpy sort."""
    from wgdi_b200.engine import ChainParams
    if case == "many_pairs":
        pts = _random_pairs_points(2600, 60000, 400, 11)
    else:
        pts = _random_pairs_points(300, 40000, 900, 12, offset=(1 << 26) - 1000)
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    blocks = gpu_engine.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
    off, perm = gpu_engine.layout()
    assert (off == np.r_[0, np.cumsum(np.bincount(pts["pair_id"], minlength=pts["n_pairs"]))]).all()
    assert (perm == np.lexsort((pts["loc2"], pts["loc1"], pts["pair_id"]))).all()          # stable: ties keep upload order
    exp = oracle_blocks(pts, prm)
    assert len(exp) == len(blocks) and len(blocks) > 100
    for b, (pair, d, l1, l2, a, sc) in enumerate(exp):
        a0, a1 = blocks.member_off[b], blocks.member_off[b + 1]
        assert blocks.pair[b] == pair and blocks.dir[b] == d and blocks.score[b] == sc and blocks.p_raw[b] == a
        assert (blocks.loc1[a0:a1] == l1).all() and (blocks.loc2[a0:a1] == l2).all()


@pytest.mark.gpu
def test_writer_field_set(gpu_engine):
    from test_host_logic import _assert_writer_fields
    from bench import make_points, params_for
    pts = make_points("C2")
    _assert_writer_fields(gpu_engine, pts, params_for(pts["cfg"]))


def test_contexts_recycle_their_blocks(gpu_engine):
    """The process-wide block cache on the device: a context built from another comparison's recycled blocks computes what
    a context on fresh cudaMalloc memory computes (full-size C2 against a C1-sized comparison in between)."""
    from test_host_logic import _assert_contexts_recycle_their_blocks
    from bench import make_points, params_for
    a, b = make_points("C2"), make_points("C1")
    _assert_contexts_recycle_their_blocks(gpu_engine.lib, a, b, params_for(a["cfg"]))
