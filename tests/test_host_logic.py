"""Host-side logic and the ABI surface, without a GPU.  Device work is served by the `emu_backend`
fixture (a g++ build of the same kernel bodies, injected by the test as a fake backend)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import PORTABLE_ENV, ROOT, check_blocks_against_golden, check_dp_against_golden, stage_e2e
from oracle import icl_oracle


def test_library_exports_every_declared_symbol():
    """The CUDA .so loads (no device needed) and exports exactly what include/icl_b200.h declares."""
    from wgdi_b200 import _lib, build
    so = build.build()
    lib = ctypes.CDLL(so)
    header = open(os.path.join(ROOT, "include", "icl_b200.h")).read()
    declared = set(re.findall(r"\b(icl_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    lib.icl_backend.restype = ctypes.c_char_p
    assert lib.icl_backend() == b"cuda"


def test_no_gpu_means_loud_failure():
    """Without a device the product path raises; it does not fall back to anything."""
    from conftest import have_gpu
    if have_gpu():
        pytest.skip("a GPU is present")
    from wgdi_b200 import _lib
    from wgdi_b200.engine import IclEngine
    _lib._LIB = None
    with pytest.raises(_lib.IclError):
        IclEngine(0)
    _lib._LIB = None


def test_package_refuses_emulation_library(tmp_path, monkeypatch):
    from wgdi_b200 import _lib, build
    so = build.build_emu(os.path.join(ROOT, "tests", "_emu"))
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", so)
    with pytest.raises(_lib.IclError):
        _lib.load()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "wgdi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libicl_oracle" not in src, f


def test_grade_hits_matches_oracle_with_ties():
    from wgdi_b200.ingest import grade_hits
    rng = np.random.default_rng(11)
    for trial in range(200):
        n, nq = int(rng.integers(1, 400)), int(rng.integers(1, 30))
        q = rng.integers(0, nq, n)
        bits = rng.integers(100, 104 + trial % 7 * 20, n).astype(np.int64 if trial % 3 else np.float64)
        mult, rn = int(rng.integers(0, 4)), int(rng.integers(1, 25))
        assert (grade_hits(q, bits, mult, rn, [50, 30, 25]) == icl_oracle.grade_hits(q, bits, mult, rn, [50, 30, 25])).all()


def test_pair_ids_follow_string_order():
    from wgdi_b200.ingest import pair_ids
    names1, names2 = ["1", "10", "2", "B2"], ["1", "2", "3"]
    c1 = np.array([0, 1, 2, 3, 1, 0]); c2 = np.array([0, 2, 1, 0, 2, 1])
    pid, names = pair_ids(names1, names2, c1, c2)
    assert names == sorted(names)
    assert names == [("1", "1"), ("1", "2"), ("10", "3"), ("2", "2"), ("B2", "1")]
    assert [names[p] for p in pid] == [(names1[a], names2[b]) for a, b in zip(c1, c2)]


def test_nargsort_desc_is_pandas():
    import pandas as pd
    from wgdi_b200.engine import nargsort_desc
    rng = np.random.default_rng(3)
    for n in (1, 5, 16, 33, 500, 5000):
        a = rng.integers(0, 9, n).astype(np.int64)
        exp = pd.DataFrame({"s": a}).sort_values(by=["s"], ascending=False).index.to_numpy()
        assert (nargsort_desc(a) == exp).all()


def test_lpt_assignment_balances():
    from wgdi_b200.shard import lpt_assign
    rng = np.random.default_rng(0)
    sizes = (rng.pareto(1.2, 441) * 1000 + 50).astype(np.int64)
    for bins in (1, 2, 4, 8):
        owner = lpt_assign(sizes, bins)
        load = np.bincount(owner, weights=sizes, minlength=bins)
        assert load.max() <= sizes.sum() / bins + sizes.max()
        assert set(owner.tolist()) <= set(range(bins))


# ---- the whole host pipeline through the fake backend -----------------------------------------------

def test_pipeline_on_golden_clouds_emu(emu_backend, clouds):
    from wgdi_b200.engine import ChainParams, IclEngine
    eng = IclEngine(0, sort_threads=1)
    try:
        for c in clouds:
            n = len(c["loc1"])
            prm = ChainParams.from_options(c["options"])
            eng.upload(np.zeros(n, np.int32), c["loc1"], c["loc2"], c["grading"], 1)
            eng.dp(prm)
            check_dp_against_golden(lambda d: (lambda s, p, dp, u: (s, p, dp, u))(*eng.debug_dp(d)), c)
            blocks = eng.extract(eng.raw_from_order(c["order1"]), eng.raw_from_order(c["order2"]))
            check_blocks_against_golden(blocks, prm.pvalue, c)
    finally:
        eng.close()


@pytest.mark.parametrize("pvalue_path", ["seq", "inc"])
def test_multi_pair_batch_matches_c_oracle_emu(emu_backend, pvalue_path, monkeypatch):
    """All chromosome pairs of a synthetic comparison in one upload vs the C oracle pair by pair, with both
    `times` accounting paths (sequential per pair / order-free incidence sort)."""
    monkeypatch.setenv("ICL_B200_PVALUE", pvalue_path)
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams, IclEngine
    from conftest import oracle_blocks
    ds = synth.make_dataset("tiny")
    pts = synth.graded_points(ds)
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    eng = IclEngine(0, sort_threads=2)
    try:
        blocks = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
        off, perm = eng.layout()
    finally:
        eng.close()
    exp = oracle_blocks(pts, prm)
    assert len(exp) == len(blocks) and len(blocks) > 20
    for b, (pair, d, l1, l2, a, sc) in enumerate(exp):
        a0, a1 = blocks.member_off[b], blocks.member_off[b + 1]
        assert blocks.pair[b] == pair and blocks.dir[b] == d and blocks.score[b] == sc
        assert (blocks.loc1[a0:a1] == l1).all() and (blocks.loc2[a0:a1] == l2).all()
        assert blocks.p_raw[b] == a


def test_forked_chains_emu(emu_backend):
    from conftest import forked_chain_points, oracle_blocks
    from wgdi_b200.engine import ChainParams, IclEngine
    pts = forked_chain_points()
    eng = IclEngine(0, sort_threads=1)
    try:
        for cov in (0.8, 0.5):
            prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25), coverage_ratio=cov)
            blocks = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], 1, prm)
            exp = oracle_blocks(pts, prm)
            assert len(blocks) == len(exp) and len(exp) >= 2
            for b, (pair, d, l1, l2, a, sc) in enumerate(exp):
                a0, a1 = blocks.member_off[b], blocks.member_off[b + 1]
                assert (blocks.loc1[a0:a1] == l1).all() and (blocks.loc2[a0:a1] == l2).all() and blocks.p_raw[b] == a
    finally:
        eng.close()


_E2E_CHILD = r"""
import sys, hashlib, json, ctypes, io, contextlib, numpy as np
sys.path.insert(0, %r)
fp = hashlib.md5(np.argsort((np.arange(1000) * 7919 %% 5).astype(np.int64), kind="quicksort").tobytes()).hexdigest()
if fp != json.load(open("meta.json"))["argsort_fingerprint"]:
    print("SKIP fingerprint", fp); sys.exit(0)
from wgdi_b200 import _lib, base
emu = %r
if emu:
    _lib._LIB = _lib._declare(ctypes.CDLL(emu))     # fake backend injected by the test
from wgdi_b200.run_colliearity import mycollinearity
import os
conf = open("icl.conf").read().replace("expected.collinearity.txt", "got.txt")
if os.environ.get("ICL_TEST_GZ"):                                     # the compressed table, as pandas would read it
    os.remove("blast.m8")
    conf = conf.replace("./blast.m8", "./blast.m8.gz")
if os.environ.get("ICL_TEST_DEVICES"):
    conf += "devices = " + os.environ["ICL_TEST_DEVICES"] + "\n"      # chromosome pairs sharded over several engines
open("icl.conf", "w").write(conf)
with contextlib.redirect_stdout(io.StringIO()):
    try:
        mycollinearity(base.load_conf("icl.conf", "collinearity")).run()
    except SystemExit as e:
        assert e.code == 0
print("OK" if open("got.txt").read() == open("expected.collinearity.txt").read() else "MISMATCH")
"""


def run_e2e_child(case, emu_path, devices=None, gz=False):
    with tempfile.TemporaryDirectory() as d:
        stage_e2e(case, d)
        env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=PORTABLE_ENV)
        if devices:
            env["ICL_TEST_DEVICES"] = devices
        if gz:
            env["ICL_TEST_GZ"] = "1"
        r = subprocess.run([sys.executable, "-c", _E2E_CHILD % (ROOT, emu_path)], cwd=d, env=env,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        if r.stdout.startswith("SKIP"):
            pytest.skip("numpy scalar-sort fingerprint differs on this machine: " + r.stdout)
        assert r.stdout.strip() == "OK", r.stdout + r.stderr


@pytest.mark.parametrize("case", ["tiny", "cross", "tiny_rev", "cross_chrom"])
def test_drop_in_driver_writes_reference_file_emu(case):
    """`mycollinearity(options).run()` (host parsing, grading, grouping, formatting) on the fake backend
    reproduces the file the unmodified reference wrote."""
    from wgdi_b200 import build
    run_e2e_child(case, build.build_emu(os.path.join(ROOT, "tests", "_emu")))


def test_drop_in_driver_reads_compressed_blast_emu():
    """`blast = table.m8.gz`: the reference reads the table with pandas, which decompresses by extension (base.py:174);
    the native parser maps raw bytes, so the driver routes compressed tables through the pandas reader — same file."""
    from wgdi_b200 import _host, build
    assert _host.is_compressed(os.path.join(ROOT, "tests", "golden", "e2e_tiny", "blast.m8.gz"))
    assert not _host.is_compressed(os.path.join(ROOT, "tests", "golden", "e2e_tiny", "genome1.gff"))
    run_e2e_child("tiny", build.build_emu(os.path.join(ROOT, "tests", "_emu")), gz=True)


def test_drop_in_driver_sharded_over_devices_emu():
    """`devices = a,b,c`: chromosome pairs LPT-sharded over three engines from one process, blocks merged on the host —
    the same file as one engine writes (the emulation ignores the device index)."""
    from wgdi_b200 import build
    run_e2e_child("cross", build.build_emu(os.path.join(ROOT, "tests", "_emu")), devices="0,0,0")


def test_cli_prints_example_and_missing_conf(capsys):
    from wgdi_b200.run import main
    assert main(["-icl", "?"]) == 0
    out = capsys.readouterr().out
    assert out.startswith("[collinearity]\ngff1 = gff1 file\n") and "positon = order" in out and "mg = 25,25" in out
    assert main(["-bi", "?"]) == 0
    out = capsys.readouterr().out
    assert out.startswith("[blockinfo]\nblast = blast file\n") and "ks_col = ks_NG86" in out
    assert main(["-icl", "/nonexistent.conf"]) == 0
    assert "not exists" in capsys.readouterr().out


# ---- streamed rank: completion flags and the sort pool -----------------------------------------------

def test_sort_pool_with_completion_flags_equals_numpy():
    """The sort helper argsorts a segment only after its flag is raised (the GPU raises them in the product,
    a thread does here) and the result is numpy's own argsort of the widened segment."""
    import threading
    import time
    from wgdi_b200 import _npsort
    rng = np.random.default_rng(7)
    sizes = rng.integers(0, 3000, 120).astype(np.int64)
    sizes[[3, 77]] = 0
    off = np.r_[0, np.cumsum(sizes)].astype(np.int64)
    keys = rng.integers(25, 400, int(off[-1])).astype(np.int32)
    order = np.argsort(-sizes, kind="stable").astype(np.int32)
    flags = np.zeros(len(sizes), np.int32)
    out = np.full(len(keys), -1, np.int32)

    def raiser():
        for s in np.argsort(sizes):                 # small segments finish first, like DP items do
            flags[s] = 1
            if s % 16 == 0:
                time.sleep(0.0005)
    th = threading.Thread(target=raiser)
    th.start()
    _npsort.seg_argsort(keys, off, order, out, 4, 0, flags, 30.0)
    th.join()
    for a, b in zip(off[:-1], off[1:]):
        assert (out[a:b] == keys[a:b].astype(np.int64).argsort(kind="quicksort")).all()
    # a flag that never comes up is an error, not a hang
    flags[:] = 0
    with pytest.raises(RuntimeError, match="timed out"):
        _npsort.seg_argsort(keys, off, order, out, 2, 0, flags, 0.2)
    # the pool survives the failed job
    _npsort.seg_argsort(keys, off, order, out, 4)
    a, b = off[5], off[6]
    assert (out[a:b] == keys[a:b].astype(np.int64).argsort(kind="quicksort")).all()


def test_dp_begin_end_contract_emu(emu_backend):
    """icl_dp_begin / icl_dp_end: same keys as icl_dp, every completion flag raised when the device is done (the
    emulation has no streaming kernel, so this is the documented degraded form), state errors reported."""
    from wgdi_b200 import synth
    from wgdi_b200._lib import IclError, icl_params
    from wgdi_b200.engine import ChainParams, IclEngine
    pts = synth.graded_points(synth.make_dataset("tiny"))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    eng = IclEngine(0, sort_threads=2)
    try:
        with pytest.raises(IclError):
            eng._check(eng.lib.icl_dp_end(eng.ctx))                     # nothing begun
        eng.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"])
        kf, kr = eng.dp(prm)
        kf, kr = kf.copy(), kr.copy()
        af, ar = eng.rank(kf, kr)
        af, ar = af.copy(), ar.copy()
        n, P = eng.n, eng.n_pairs
        keys = np.zeros(2 * n, np.int32)
        flags = np.full(2 * P, -7, np.int32)
        p = icl_params(25, 25, 50, 3, 0.8)
        eng._check(eng.lib.icl_dp_begin(eng.ctx, ctypes.byref(p), keys.ctypes.data, keys.ctypes.data + 4 * n, flags.ctypes.data))
        with pytest.raises(IclError):                                    # extraction needs icl_dp_end first
            eng.extract(af, ar)
        eng._check(eng.lib.icl_dp_end(eng.ctx))
        assert (flags == 1).all()
        assert (keys[:n] == kf).all() and (keys[n:] == kr).all()
        sf, sr = eng.dp_rank_streamed(prm)
        assert (sf == af).all() and (sr == ar).all()
        b1 = eng.extract(sf, sr)
        b2 = eng.extract(af, ar)
        assert len(b1) == len(b2) and (b1.p_raw == b2.p_raw).all() and (b1.loc1 == b2.loc1).all()
    finally:
        eng.close()


def test_grade_hits_chromosomes_matches_oracle():
    """`comparison = chromosomes` grading (run_colliearity.py:42-57): vectorised host code vs the oracle's loop, with
    tied bit scores, several chromosome pairs and gene names whose string order differs from their numeric order."""
    from wgdi_b200.ingest import grade_hits_chromosomes
    rng = np.random.default_rng(3)
    n_genes, n = 300, 6000
    names = np.array([f"g{i}" for i in rng.permutation(n_genes)])          # "g10" < "g9": string order
    g1 = rng.integers(0, n_genes, n)
    g2 = rng.integers(0, 150, n)
    pair = rng.integers(0, 7, n)
    bits = rng.integers(100, 130, n).astype(np.int64)
    name_rank = np.empty(n_genes, np.int64)
    name_rank[np.argsort(names, kind="stable")] = np.arange(n_genes)
    for multiple, rn, grading in ((1, 10, (50, 30, 25)), (2, 5, (60, 35, 20)), (3, 4, (9, 5, 1))):
        ga, oa = grade_hits_chromosomes(name_rank, g1, g2, pair, bits, multiple, rn, grading)
        gb, ob = icl_oracle.grade_hits_chromosomes(names[g1], g2, pair, bits, multiple, rn, grading)
        assert (ga == gb).all() and (oa == ob).all()
        assert set(np.unique(ga)) <= {0, *grading} and (ga > 0).any() and (ga == 0).any()
    for bad in ((0, 10), (10, 10), (12, 10)):                              # pd.cut: bins must increase
        with pytest.raises(ValueError):
            grade_hits_chromosomes(name_rank, g1, g2, pair, bits, bad[0], bad[1], (50, 30, 25))
        with pytest.raises(ValueError):
            icl_oracle.grade_hits_chromosomes(names[g1], g2, pair, bits, bad[0], bad[1], (50, 30, 25))


def test_batch_runner_emu(emu_backend):
    """BatchRunner (two engines, two host threads, shared sort pool) returns every job's blocks in job order,
    identical to a lone engine fed the points RE-GRADED for that job's `grading` (run_colliearity.py:64-86: a grading
    sweep changes every point's value); a failing job raises in the caller."""
    from conftest import oracle_blocks
    from wgdi_b200 import synth
    from wgdi_b200.batch import BatchRunner, regrade, sweep_jobs
    from wgdi_b200.engine import ChainParams, IclEngine
    ds = synth.make_dataset("tiny")
    pts = synth.graded_points(ds)
    gradings = [(50, 30, 25), (50, 40, 25), (40, 40, 10)]
    jobs, keys = sweep_jobs(pts, gradings, [(25, 25), (40, 40), (10, 15)], [1])
    assert all("grading" not in p_ for p_, _ in jobs)       # class ids travel, values come from the job
    runner = BatchRunner(0, depth=2, sort_threads=2)
    try:
        got = runner.run(jobs)
        bad = dict(jobs[0][0], loc1=pts["loc1"][:-1])
        with pytest.raises(ValueError):
            runner.run(jobs[:1] + [(bad, jobs[0][1])])
    finally:
        runner.close()
    with pytest.raises(ValueError):
        sweep_jobs({k: v for k, v in pts.items() if k != "grade_class"}, gradings, [(25, 25)], [1])
    eng = IclEngine(0, sort_threads=2)
    try:
        for (p_, prm), (g, mg, _), b in zip(jobs, keys, got):
            own = synth.graded_points(ds, grading=g)         # graded from scratch with this job's values
            assert (own["grading"] == regrade(pts, g)["grading"]).all()
            exp = eng.chain(own["pair_id"], own["loc1"], own["loc2"], own["grading"], own["n_pairs"], prm)
            assert len(b) == len(exp) > 0 and (b.p_raw == exp.p_raw).all() and (b.loc1 == exp.loc1).all()
            assert (b.score == exp.score).all()
        # the oracle on independently re-graded input, one job per grading
        for j in (0, 3, 6):
            own = synth.graded_points(ds, grading=keys[j][0])
            exp = oracle_blocks(own, jobs[j][1])
            assert len(exp) == len(got[j])
            for bi, (pair, d, l1, l2, a, sc) in enumerate(exp):
                a0, a1 = got[j].member_off[bi], got[j].member_off[bi + 1]
                assert got[j].score[bi] == sc and got[j].p_raw[bi] == a and (got[j].loc1[a0:a1] == l1).all()
    finally:
        eng.close()
    # the grading sweep really changes the result (same mg, gradings 50,30,25 vs 50,40,25)
    assert len(got[0]) != len(got[3]) or (got[0].score[:50] != got[3].score[:50]).any()
    assert len({len(b) for b in got}) > 1                  # the mg sweep really changes the result


def test_device_rank_mode_emu(emu_backend):
    """icl_rank_device (non-reference option): blocks equal the oracle's under a stable descending sort; same block
    set as the reference order wherever no top scores tie; state errors reported."""
    from conftest import oracle_blocks_stable
    from wgdi_b200 import synth
    from wgdi_b200._lib import IclError
    from wgdi_b200.engine import ChainParams, IclEngine
    pts = synth.graded_points(synth.make_dataset("tiny"))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    eng = IclEngine(0, sort_threads=2)
    try:
        eng.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"])
        with pytest.raises(IclError):
            eng._check(eng.lib.icl_rank_device(eng.ctx, 0))              # before icl_dp
        got = eng.run_resident_device_rank(prm)
        ref = eng.run_resident(prm)
    finally:
        eng.close()
    exp = oracle_blocks_stable(pts, prm)
    assert len(got) == len(exp) > 20
    for b, (pair, d, l1, l2, a, sc) in enumerate(exp):
        a0, a1 = got.member_off[b], got.member_off[b + 1]
        assert got.pair[b] == pair and got.dir[b] == d and got.score[b] == sc and got.p_raw[b] == a
        assert (got.loc1[a0:a1] == l1).all() and (got.loc2[a0:a1] == l2).all()
    assert sorted(got.score.tolist()) == sorted(ref.score.tolist()) or len(got) != len(ref)   # tie order may differ, nothing else


def test_device_rank_hands_back_to_the_host_emu(emu_backend):
    """When the device cannot rank a comparison with the requested numpy flavour (ICL_ERR_STATE: on a GPU the std::sort
    heap-sort branch, here the emulated backend, which has no such kernels), the engine ranks that comparison with numpy
    on the host: same blocks as the host-ranked path, the hand-back counted; the start-up probe finds no flavour."""
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams, IclEngine
    pts = synth.graded_points(synth.make_dataset("tiny"))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    eng = IclEngine(0, sort_threads=2)
    try:
        eng.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"])
        eng.rank_mode = "host"
        ref = eng.run_resident(prm)
        for k, mode in enumerate(("numpy_avx512", "numpy_avx2", "numpy_scalar")):
            got = eng.run_resident_device_rank(prm, mode=mode)
            assert eng.host_rank_fallbacks == k + 1
            assert len(got) == len(ref) > 20
            for f in ("pair", "dir", "score", "m", "N", "N1", "p_raw", "member_off", "loc1", "loc2", "member_idx"):
                assert (getattr(got, f) == getattr(ref, f)).all(), (mode, f)
        assert eng.numpy_sort_kind() is None
        eng.rank_mode = "auto"
        auto = eng.run_resident(prm)                                     # auto with no matching flavour = the host path
        assert (auto.score == ref.score).all() and (auto.loc1 == ref.loc1).all()
    finally:
        eng.close()


def test_avx512_argsort_specification_matches_numpy():
    """tools/xss_emulation.py — the plain-Python statement of numpy's AVX-512 argsort that the CUDA kernel follows —
    against numpy itself, where numpy sorts that way (AVX-512 hosts); elsewhere there is nothing to compare with."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import xss_emulation
    if not xss_emulation.is_avx512_numpy():
        pytest.skip("numpy on this machine does not dispatch int64 argsort to its AVX-512 kernels")
    assert xss_emulation.self_check(cases=120, seed=7, verbose=False) == 0


def test_no_progress_passes_in_closed_form():
    """The closed-form permutation of a partition pass that moves everything to one side (and runs of such passes) —
    what the CUDA kernel applies instead of replaying them — against the sequential specification."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import xss_emulation
    assert xss_emulation.check_fast_forward(cases=24)


def test_partition_in_merge_order_equals_the_sequential_partition():
    """The closed form behind the rank kernel's CTA-cooperative phase: the groups of a partition pass are taken in the
    order of a merge of two prefix-sum sequences, so every group's store pointers can be computed independently."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import xss_emulation as x
    assert x.check_merge_path(trials=200) == 0
    assert x.check_std_sort_equal() == 0          # std::sort on equal keys is a fixed permutation (xs_std_equal_pos)


def test_avx2_argsort_specification_matches_numpy():
    """The 4-lane flavour of the same specification against numpy in a child process with every AVX512* feature
    disabled (numpy then dispatches int64 argsort to its AVX2 kernels, where the CPU has AVX2)."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); import xss_emulation as x\n"
            "if not x.is_avx2_numpy(): print('SKIP'); raise SystemExit(0)\n"
            "print('BAD', x.self_check(cases=120, seed=7, verbose=False, fn=x.avx2_argsort))" % os.path.join(ROOT, "tools"))
    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES="AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX512_KNL AVX512_KNM")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    if "SKIP" in r.stdout:
        pytest.skip("numpy on this machine has no AVX2 argsort to compare with")
    assert "BAD 0" in r.stdout, r.stdout


def test_place_blocks_equals_merge_blocks():
    """The parallel host gather of the strong-scaling path (every rank writes its blocks to their place in the merged arrays,
    shard.place_blocks) gives exactly what the sequential merge (shard.merge_blocks) gives."""
    from wgdi_b200 import shard
    from wgdi_b200.engine import Blocks
    rng = np.random.default_rng(4)
    P, world = 23, 3
    owner = rng.integers(0, world, P)
    parts, gps, rowss = [], [], []
    for r in range(world):
        mine = np.flatnonzero(owner == r)
        nbp = rng.integers(0, 6, len(mine))                       # blocks per local pair (some pairs without blocks)
        pair = np.repeat(np.arange(len(mine)), nbp).astype(np.int32)
        nb = len(pair)
        m = rng.integers(3, 9, nb).astype(np.int32)
        nm = int(m.sum())
        rows = rng.permutation(5000)[:4000]
        parts.append(Blocks(pair, rng.integers(0, 2, nb).astype(np.int32), rng.integers(0, 999, nb).astype(np.int64), m,
                            m + 1, m + 2, (m + 3).astype(np.int64), m + 4, m + 5, rng.random(nb),
                            np.r_[0, np.cumsum(m)].astype(np.int64), rng.integers(0, 99, nm).astype(np.int32),
                            rng.integers(0, 99, nm).astype(np.int32), rng.integers(0, 4000, nm).astype(np.int32)))
        gps.append(mine)
        rowss.append(rows)
    # sequential merge on globally renumbered copies
    glob = []
    for b, gp, rows in zip(parts, gps, rowss):
        g = Blocks(*[getattr(b, f) for f in shard.BLOCK_FIELDS])
        g.pair = gp[b.pair].astype(np.int32)
        g.member_idx = rows[b.member_idx].astype(np.int32)
        glob.append(g)
    exp = shard.merge_blocks(glob)
    cb = sum(shard.pair_counts(b, gp, P)[0] for b, gp in zip(parts, gps))
    cm = sum(shard.pair_counts(b, gp, P)[1] for b, gp in zip(parts, gps))
    bb, mb = np.r_[0, np.cumsum(cb)], np.r_[0, np.cumsum(cm)]
    out = {f: np.zeros(int(mb[-1]) if f in shard.MEMBER_FIELDS else int(bb[-1]) + 1, getattr(exp, f).dtype) for f in shard.BLOCK_FIELDS}
    for b, gp, rows in zip(parts, gps, rowss):
        shard.place_blocks(b, gp, rows, bb, mb, out)
    out["member_off"][int(bb[-1])] = int(mb[-1])
    for f in shard.BLOCK_FIELDS:
        n = len(getattr(exp, f))
        assert (out[f][:n] == getattr(exp, f)).all(), f


def _assert_selection_is_a_filter(eng, pts, prm, min_members):
    """The blocks returned for `min_members` are exactly the full result with the smaller blocks dropped — same order,
    same fields (N / N1 / p still count every accepted block, collinearity.py:150), same gene pairs."""
    full = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm, min_members=1)
    part = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm, min_members=min_members)
    keep = np.flatnonzero(full.m >= min_members)
    assert 0 < len(keep) < len(full) and len(part) == len(keep)
    for f in ("pair", "dir", "score", "m", "L", "N", "N1", "L1", "L2", "p_raw"):
        assert (getattr(part, f) == getattr(full, f)[keep]).all(), f
    assert (np.diff(part.member_off) == part.m).all() and part.member_off[0] == 0
    sel = np.concatenate([np.arange(full.member_off[b], full.member_off[b + 1]) for b in keep])
    for f in ("loc1", "loc2", "member_idx"):
        assert (getattr(part, f) == getattr(full, f)[sel]).all(), f
    # emission order: pair ascending, forward before anti-diagonal inside a pair (run_colliearity.py:118, collinearity.py:49-68)
    order = part.pair.astype(np.int64) * 2 + part.dir
    assert (np.diff(order) >= 0).all()


def test_min_members_selection_is_a_filter_emu(emu_backend):
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams, IclEngine
    ds = synth.make_dataset("tiny")
    pts = synth.graded_points(ds)
    eng = IclEngine(0, sort_threads=2)
    try:
        for mm in (4, 5, 9):
            _assert_selection_is_a_filter(eng, pts, ChainParams(mg1=25, mg2=25, grading=(50, 30, 25)), mm)
    finally:
        eng.close()


@pytest.mark.parametrize("case", ["many_pairs", "wide_keys"])
def test_upload_layout_paths_emu(emu_backend, case):
    """Host logic of the layout planned at upload (pair offsets, dense row ranges, item order) with many pairs / 64-bit keys."""
    from conftest import oracle_blocks
    from test_gpu_parity import _random_pairs_points
    from wgdi_b200.engine import ChainParams, IclEngine
    pts = _random_pairs_points(2600, 20000, 300, 21) if case == "many_pairs" else _random_pairs_points(40, 8000, 500, 22, offset=(1 << 26) - 600)
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    eng = IclEngine(0, sort_threads=2)
    try:
        blocks = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm)
        off, perm = eng.layout()
    finally:
        eng.close()
    assert (off == np.r_[0, np.cumsum(np.bincount(pts["pair_id"], minlength=pts["n_pairs"]))]).all()
    assert (perm == np.lexsort((pts["loc2"], pts["loc1"], pts["pair_id"]))).all()
    exp = oracle_blocks(pts, prm)
    assert len(exp) == len(blocks) and len(blocks) > 50
    for b, (pair, d, l1, l2, a, sc) in enumerate(exp):
        a0, a1 = blocks.member_off[b], blocks.member_off[b + 1]
        assert blocks.pair[b] == pair and blocks.dir[b] == d and blocks.score[b] == sc and blocks.p_raw[b] == a
        assert (blocks.loc1[a0:a1] == l1).all() and (blocks.loc2[a0:a1] == l2).all()


def _assert_writer_fields(eng, pts, prm):
    """icl_extract_fields(ICL_FIELDS_WRITER): the arrays the file is written from are those of the full result, the others
    do not come back."""
    eng.fields = "all"
    full = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm, min_members=5)
    eng.fields = "writer"
    try:
        lean = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm, min_members=5)
    finally:
        eng.fields = "all"
    assert len(lean) == len(full) > 0
    for f in ("pair", "score", "m", "p_raw", "member_off", "loc1", "loc2"):
        assert (getattr(lean, f) == getattr(full, f)).all(), f
    for f in ("dir", "L", "N", "N1", "L1", "L2", "member_idx"):
        assert getattr(lean, f) is None and getattr(full, f) is not None, f


def test_writer_field_set_emu(emu_backend):
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams, IclEngine
    pts = synth.graded_points(synth.make_dataset("tiny"))
    eng = IclEngine(0, sort_threads=2)
    try:
        _assert_writer_fields(eng, pts, ChainParams(mg1=25, mg2=25, grading=(50, 30, 25)))
    finally:
        eng.close()


def _assert_contexts_recycle_their_blocks(lib, pts_a, pts_b, prm):
    """Contexts created one after the other (what the drop-in driver, the batch runner and the sharded runs do): the
    second one is served from the process-wide block cache (include/icl_b200.h, icl_cache_stats) — recycled device blocks
    arrive zeroed — and computes exactly what a fresh process would; icl_cache_trim hands everything back."""
    from wgdi_b200.engine import IclEngine

    def stats():
        out = np.zeros(3, np.int64)
        assert lib.icl_cache_stats(out.ctypes.data, 3) == 3
        return out.tolist()

    def run(pts):
        eng = IclEngine(0, sort_threads=2)
        try:
            b = eng.chain(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"], prm, min_members=1)
            return {k: np.array(getattr(b, k), copy=True) for k in ("pair", "dir", "score", "m", "N", "N1", "p_raw", "member_off",
                                                                    "loc1", "loc2")}
        finally:
            eng.close()
    lib.icl_cache_trim()
    assert stats()[2] == 0
    first = run(pts_a)
    h0, m0, held = stats()
    assert held > 0                                          # the closed context's blocks are kept
    other = run(pts_b)                                       # another comparison dirties the recycled blocks
    again = run(pts_a)
    h1, m1, _ = stats()
    assert h1 > h0                                           # ... and they were reused
    for k in first:
        assert np.array_equal(first[k], again[k]), k
    assert len(other["pair"]) > 0
    lib.icl_cache_trim()
    assert stats()[2] == 0
    fresh = run(pts_a)                                       # straight from the driver again
    for k in first:
        assert np.array_equal(first[k], fresh[k]), k


def test_contexts_recycle_their_blocks_emu(emu_backend):
    from wgdi_b200 import synth
    from wgdi_b200.engine import ChainParams
    a = synth.graded_points(synth.make_dataset("tiny"))
    b = synth.graded_points(synth.make_dataset("cross"))
    _assert_contexts_recycle_their_blocks(emu_backend, a, b, ChainParams(mg1=25, mg2=25, grading=(50, 30, 25)))


def test_compact_points_with_known_bounds():
    """`compact_points` narrows by the bounds the ingest knows instead of scanning: same arrays whenever the bounds are the
    real extremes, a wider (never a narrower) element type when they are only an upper bound."""
    from wgdi_b200.params import compact_points
    from wgdi_b200.run_colliearity import _point_bounds
    rng = np.random.default_rng(2)
    pair, l1, l2, g = rng.integers(0, 300, 1000), rng.integers(1, 200, 1000), rng.integers(1, 70000, 1000), rng.choice([50, 30, 25], 1000)
    plain = compact_points(pair, l1, l2, g)
    exact = compact_points(pair, l1, l2, g, ((pair.min(), pair.max()), (l1.min(), l1.max()), (l2.min(), l2.max()), (25, 50)))
    assert [a.dtype for a in plain] == [a.dtype for a in exact] == [np.uint16, np.uint8, np.int32, np.uint8]
    assert all((a == b).all() for a, b in zip(plain, exact))
    genes1, genes2 = dict(order=np.arange(1, 1001)), dict(order=np.arange(1, 201))
    b = _point_bounds(300, genes1, genes2, [50, 30, 25])
    assert b == ((0, 299), (1, 1000), (1, 200), (25, 50))
    wide = compact_points(pair, l1, np.minimum(l2, 200), g, b)
    assert [a.dtype for a in wide] == [np.uint16, np.uint16, np.uint8, np.uint8] and (wide[1] == l1).all()
    neg = compact_points(pair, l1, l2, g, ((0, 299), (-5, 10), (1, 70000), (25, 50)))
    assert neg[1].dtype == np.int32                     # a possible negative value stays int32 (the upload rejects it)
    assert all(len(a) == 0 and a.dtype == np.int32 for a in compact_points([], [], [], [], b))


def test_drop_in_driver_stops_like_the_reference_when_nothing_survives_emu(emu_backend, tmp_path, capsys):
    """No hit passes the filters: the reference prints its message and exits with 1 (run_colliearity.py:113-115).  The
    engine that the driver had started in the background while the files were read is closed, nothing is written."""
    import contextlib
    import io
    from conftest import stage_e2e
    from wgdi_b200 import base, engine
    from wgdi_b200.run_colliearity import mycollinearity
    stage_e2e("tiny", str(tmp_path))
    closed = []
    real_close = engine.IclEngine.close
    def spy(self):
        closed.append(self)
        return real_close(self)
    engine.IclEngine.close = spy
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for mode in ("native", "native-steps", "pandas"):
            closed.clear()
            opts = [(k, "1000000" if k == "score" else "none.txt" if k == "savefile" else v)
                    for k, v in base.load_conf("icl.conf", "collinearity")] + [("ingest", mode)]
            out = io.StringIO()
            with contextlib.redirect_stdout(out), pytest.raises(SystemExit) as e:
                mycollinearity(opts).run()
            assert e.value.code == 1 and "Stopped!" in out.getvalue(), mode
            assert len(closed) >= 1 and all(c.ctx is None for c in closed), mode
            assert not os.path.exists("none.txt")
    finally:
        os.chdir(cwd)
        engine.IclEngine.close = real_close


def test_classes_of_ambiguous_fast_path_equals_the_loop(monkeypatch):
    """The one-call ranking of the tie-ambiguous queries (numpy's own int64 argsort routine through the sort helper) gives
    what the per-query loop over `nargsort_desc` gives; float bit scores and huge values take the loop."""
    from wgdi_b200 import engine, ingest
    if not engine._npsort_ok():
        pytest.skip("_npsort helper not built")
    rng = np.random.default_rng(8)
    for trial in range(30):
        lens = rng.integers(1, 260, int(rng.integers(1, 40)))
        off = np.r_[0, np.cumsum(lens)].astype(np.int64)
        bits = rng.integers(100, 100 + int(rng.integers(2, 30)), int(off[-1])).astype(np.int64)
        mult, rn = int(rng.integers(1, 4)), int(rng.integers(2, 40))
        fast = ingest.classes_of_ambiguous(off, bits, mult, rn)
        with monkeypatch.context() as m:
            m.setattr(engine, "_npsort_ok", lambda: False)
            slow = ingest.classes_of_ambiguous(off, bits, mult, rn)
        assert (fast == slow).all()
        halves = ingest.classes_of_ambiguous(off, bits.astype(np.float64) / 2, mult, rn)      # float scores: the loop
        assert sorted(halves.tolist()) == sorted(slow.tolist())                # same class sizes (tie order is per dtype)
    assert len(ingest.classes_of_ambiguous(np.zeros(1, np.int64), np.zeros(0, np.int64), 1, 5)) == 0
