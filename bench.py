#!/usr/bin/env python
"""Benchmark of the `wgdi -icl` hot path (BASELINE.json metric: BLAST hits chained per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2]

A "step" is one pass of the hot path over one batch of synthetic input: all chromosome pairs of
BASELINE.json config 2 (Vitis-sized self comparison, 19 chromosomes -> 361 chromosome pairs, ~2M graded
BLAST hits) are sorted, chained in both directions, ranked in the order numpy's own unstable argsort would give
(required for the reference's tie order) and extracted into blocks with p values.  The rank runs on the device when one
of its exact reproductions of numpy's argsort (AVX-512 / AVX2 / scalar) matches numpy in this process — checked at
start-up, `config.tie_order` reports it — and with numpy on the host otherwise.

  value   hits/s with the graded hits already resident in HBM (icl_dp + rank + icl_extract)
  e2e     hits/s through the engine's public call with HOST buffers: pinned inputs -> H2D -> ... -> blocks on host
  roofline  the chaining-DP kernel against the measured HBM bandwidth (it is latency-bound: DESIGN.md)
  cpu_baseline  the oracle's C restatement of the reference algorithm on all host cores (kind "port")
  host_ranked / portable_numpy / avx2_numpy / device_rank_option / batch / icl_wall   comparison legs: the same step
          ranked by numpy on the host; the same in child processes whose numpy runs its scalar resp. AVX2 argsort (device-
          ranked with that flavour, blocks checked against the oracle in the child); the stable-rank option (NOT the
          reference's tie order); pipelined comparisons; conf-in -> file-out wall time

With N > 1 (torchrun, one rank per GPU) the workload is N such comparisons (seeds differ); their 361*N chromosome
pairs are pooled and assigned to ranks by longest-processing-time on point count; every rank uploads only its
pairs; there is no data-path collective (scaling: weak).  NCCL is used for the barrier and the max-over-ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DP_BYTES_PER_POINT = 42          # 2 directions x [12 B read + 9 B written]
EXPORT_BYTES_PER_POINT = 4       # streamed export of the DP kernel: the mirrored direction reads its position map (4 B);
                                 # the 8 B/point of score writes go to page-locked host memory over PCIe and are not counted
METRIC = "blast_hits_chained_per_sec"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the DP kernel on C2 (ncu --set full captures): without the
# export (device-ranked, the default: profiles/r02_chain_ring_ncu_summary.md) and with the streamed export to the host
# (host-ranked: profiles/r01_chain_ring_ncu_summary.md)
NCU_TRAFFIC_C2 = {"dp_only": 72.0e6, "streamed_export": 81.3e6}


WORKLOADS = {   # BASELINE.json `configs`, as built by wgdi_b200/synth.py
    "C1": "BASELINE.json config 1 (5-chromosome self comparison, ~1M graded hits, 25 chromosome pairs)",
    "C2": "BASELINE.json config 2 (Vitis-sized self comparison, 19 chromosomes, ~2M graded hits, 361 chromosome pairs)",
    "C3": "BASELINE.json config 3 (cross-species, two 12-chromosome genomes, ~5M graded hits, 144 chromosome pairs)",
    "C4": "BASELINE.json config 4 (hexaploid-wheat-sized all-vs-all, 21x21 chromosome pairs, ~10M graded hits)",
}


def workload_label(name):
    return f"{name}: {WORKLOADS.get(name, 'synthetic workload ' + name)}"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.rows.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 6 or not c[0].isdigit():
                continue
            sm.append(int(c[0])); mx.append(int(c[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_points(workload, seed_shift=0, scale=1.0):
    from dataclasses import replace
    from wgdi_b200 import synth
    cfg = synth.CONFIGS[workload]
    if seed_shift:
        cfg = replace(cfg, seed=cfg.seed + 1000 * seed_shift)
    ds = synth.make_dataset(cfg, scale=scale)
    pts = synth.graded_points(ds)
    pts["cfg"] = ds.cfg
    return pts


def params_for(cfg):
    from wgdi_b200.params import ChainParams            # plain dataclass: the CPU arm loads nothing of the CUDA product
    return ChainParams(mg1=cfg.mg[0], mg2=cfg.mg[1], grading=tuple(cfg.grading), pvalue=cfg.pvalue)


def cpu_port_pass(pts, prm, threads):
    """One pass of the oracle's C restatement (chain both directions, numpy rank, extraction) on `threads` cores."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle, icl_oracle
    o = np.lexsort((pts["loc2"], pts["loc1"], pts["pair_id"]))
    off = np.r_[0, np.cumsum(np.bincount(pts["pair_id"], minlength=pts["n_pairs"]))].astype(np.int64)
    op = icl_oracle.IclParams([("grading", ",".join(map(str, prm.grading))), ("mg", f"{prm.mg1},{prm.mg2}")])
    l1, l2, g = pts["loc1"][o], pts["loc2"][o], pts["grading"][o]
    pool = ThreadPoolExecutor(threads) if threads > 1 else None
    t0 = time.perf_counter()
    bt = c_oracle.Batch(off, l1, l2, g, op, threads=threads)
    bt.chain()
    bt.rank(pool)
    bt.extract()
    dt = time.perf_counter() - t0
    if pool:
        pool.shutdown()
    return dt, int(bt.n_blk[0].sum() + bt.n_blk[1].sum())


def icl_wall_time(workload, device, bi=True, repeats=3):
    """`-icl` wall time per genome pair: conf in -> collinearity file out through the drop-in driver
    (`mycollinearity(options).run()`: parse gff/lens/m8, filter, grade, GPU chaining, write the file)."""
    import contextlib
    import io
    import shutil
    import tempfile
    from wgdi_b200 import base, synth
    from wgdi_b200.run_colliearity import mycollinearity
    root = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="icl_wall_", dir=root)
    try:
        ds = synth.make_dataset(workload)
        conf = synth.write_dataset(ds, d)
        blast_mb = os.path.getsize(os.path.join(d, "blast.m8")) / 1e6
        best, timing, native, runs = None, None, None, []
        for _ in range(repeats):
            opts = list(base.load_conf(conf, "collinearity")) + [("device", str(device))]
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                m = mycollinearity(opts)
                try:
                    m.run()
                except SystemExit:
                    pass
            dt = time.perf_counter() - t0
            runs.append(dt)
            if best is None or dt < best:
                best, timing, native = dt, m.timing, m.native
        out_mb = os.path.getsize(dict(opts)["savefile"]) / 1e6
        res = {"value": best, "unit": "s", "what": f"{workload}: conf in -> collinearity file out, best of {repeats}, files in tmpfs",
               "blast_rows": int(len(ds.q)), "blast_mb": round(blast_mb, 1), "out_mb": round(out_mb, 1),
               "ingest_s": timing.get("ingest_s"), "chain_s": timing.get("chain_s"), "write_s": timing.get("write_s"),
               # every run builds its own context; from the second context of a process on its device arrays come from the
               # library's block cache (icl_cache_stats), the first asks the driver for them
               "runs_s": [round(x, 4) for x in runs],
               "host_path": "native (libicl_host.so)" if native else "pandas"}
        if not bi:
            return res
        # the tool every user runs next: `-bi` on the file just written, conf in -> CSV out
        try:
            from wgdi_b200.block_info import block_info
            o = dict(opts)
            bi = [(k, o[k]) for k in ("blast", "gff1", "gff2", "lens1", "lens2")] + [
                ("collinearity", o["savefile"]), ("score", "100"), ("evalue", "1e-5"), ("repeat_number", "20"),
                ("position", "order"), ("ks", "none"), ("ks_col", "ks_NG86"), ("savefile", os.path.join(d, "blockinfo.csv")),
                ("device", str(device))]
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                b = block_info(bi)
                b.run()
            res["bi_wall"] = {"value": time.perf_counter() - t0, "unit": "s", "what": "-bi conf in -> block-information CSV out",
                              "blocks": b.timing.get("blocks"), "ingest_s": b.timing.get("ingest_s"),
                              "device_s": b.timing.get("device_s"), "write_s": b.timing.get("write_s")}
        except Exception as e:                         # the -icl numbers above stand on their own
            res["bi_wall"] = {"error": f"{type(e).__name__}: {e}"}
        return res
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_sample(pts, prm, k, max_points, gpu_blocks=None):
    """The REAL reference (unmodified wgdi/collinearity.py from oracle/_ref, see oracle/make_ref.py) on a stratified sample
    of chromosome pairs at the inner seam `collinearity(options, points).run()`, one pair per host core; the whole-workload
    figure is an extrapolation with the per-pair model fitted on the sample.  With `gpu_blocks` (all blocks
    of >= 5 gene pairs of the same comparison from the CUDA path) the sampled pairs are compared block by block."""
    from oracle import ref_leg
    if not ref_leg.available():
        return {"unavailable": "oracle/_ref/wgdi/collinearity.py not present (run `make -C oracle ref` where /root/reference exists)"}
    options = [("grading", ",".join(map(str, prm.grading))), ("mg", f"{prm.mg1},{prm.mg2}"), ("pvalue", str(prm.pvalue))]
    r = ref_leg.time_sample(pts, options, k=k, max_points=max_points)
    blocks = r.pop("blocks")
    out = {"kind": "reference", "what": "unmodified wgdi/collinearity.py: collinearity(options, points).run() per chromosome pair, "
           "one pair per process; whole-workload numbers are EXTRAPOLATED with t = a*n + b*n^2 fitted on the sample and the "
           "reference's static chunking of pairs over `cores` processes (run_colliearity.py:123-132)",
           "sample": f"{len(r['pairs'])} chromosome pairs of the workload (largest" + (f" <= {max_points} points" if max_points else "") +
                     ", median, smallest, random), " + f"{sum(r['points'])} points, measured",
           **r}
    if gpu_blocks is not None:
        same, checked = True, 0
        for p, ref_blocks in blocks.items():
            keep = [b for b in ref_blocks if len(b[0]) >= 5]                       # run_colliearity.py:178
            idx = np.flatnonzero(gpu_blocks.pair == p)
            pv = np.round(gpu_blocks.p_raw[idx], 4)                                 # collinearity.py:153
            idx = idx[~(pv > prm.pvalue)]
            if len(idx) != len(keep):
                same = False
                continue
            for b, (l1, l2, pval, sc) in zip(idx, keep):
                a0, a1 = gpu_blocks.member_off[b], gpu_blocks.member_off[b + 1]
                same &= bool((gpu_blocks.loc1[a0:a1] == l1).all() and (gpu_blocks.loc2[a0:a1] == l2).all() and
                             gpu_blocks.score[b] == sc and np.round(gpu_blocks.p_raw[b], 4) == pval)
                checked += 1
        out["blocks_compared_with_gpu"] = checked
        out["identical_to_gpu"] = bool(same)
    return out


def run_reference_arm(args, rank, world):
    """The reference's algorithm on the host cores.  The reference is pure Python (6-8 ms per point per core,
    BASELINE.md): the timed steps run the oracle's C port of it (kind "port"), multi-threaded over chromosome pairs like the
    reference's worker pool, on the whole workload; next to it `cpu_baseline.reference_sample` times the unmodified
    reference itself on a stratified sample of chromosome pairs (oracle/_ref) and extrapolates.  Nothing of the CUDA
    product is imported here."""
    if rank != 0:
        return
    from oracle import c_oracle
    c_oracle.build()
    pts = make_points(args.workload)
    prm = params_for(pts["cfg"])
    threads = os.cpu_count() or 1
    n = len(pts["loc1"])
    for _ in range(args.warmup):
        cpu_port_pass(pts, prm, threads)
    times = [cpu_port_pass(pts, prm, threads)[0] for _ in range(args.steps)]
    T = float(np.sum(times))
    value = n * args.steps / T
    sample = None
    if not args.no_reference_sample:
        try:
            sample = reference_sample(pts, prm, k=min(8, threads), max_points=None)
        except Exception as e:                              # noqa: BLE001 — the timed port steps stand on their own
            sample = {"error": f"{type(e).__name__}: {e}"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "hits/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_label(args.workload), "graded_hits": n, "chromosome_pairs": pts["n_pairs"],
                   "note": "one comparison on the host cores whatever --gpus says: at N > 1 the GPU arm runs N such comparisons, "
                           "so the driver's ratio compares RATES (hits/s), not equal amounts of work"},
        "cpu_baseline": {"value": value, "unit": "hits/s", "cores": threads, "kind": "port",
                         "sample": f"full {args.workload} workload per step ({n} hits), C port of the reference algorithm",
                         "reference_sample": sample},
        "e2e": {"value": value, "unit": "hits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


PORTABLE_NUMPY = ("AVX2 FMA3 AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX512_KNL AVX512_KNM")


def portable_child(args):
    """Runs in a child process whose numpy has its SIMD sorts disabled (NPY_DISABLE_CPU_FEATURES, set by the parent before
    numpy is imported): numpy.argsort is then its scalar introsort, which the device reproduces exactly, so the engine
    ranks on the GPU and the host is out of the step — still the reference's tie order for THAT numpy.  Prints one JSON
    object: step time with inputs resident, and whether the blocks equal the C oracle's ranked by this process's numpy."""
    import torch
    from wgdi_b200.engine import IclEngine
    dev = args.portable_child
    torch.cuda.set_device(dev)
    pts = make_points(args.workload, seed_shift=args.seed_shift)
    from wgdi_b200 import synth
    prm = params_for(synth.CONFIGS[args.workload])
    eng = IclEngine(dev)
    scalar = eng.numpy_sort_is_scalar()
    eng.upload(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"], pts["n_pairs"])
    for _ in range(max(args.warmup, 3)):
        blocks = eng.run_resident(prm, min_members=5, copy=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    t = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        blocks = eng.run_resident(prm, min_members=5, copy=False)
        torch.cuda.synchronize()
        t += time.perf_counter() - t0
    # end to end: page-locked host inputs -> H2D -> ... -> blocks on the host
    from wgdi_b200.params import compact_points
    host_in = {}
    for k, arr in zip(("pair_id", "loc1", "loc2", "grading"), compact_points(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"])):
        buf = eng._buf("in_" + k, len(arr), arr.dtype)
        buf[:] = arr
        host_in[k] = buf
    te = 0.0
    for i in range(args.steps + 2):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        blocks = eng.chain(host_in["pair_id"], host_in["loc1"], host_in["loc2"], host_in["grading"], pts["n_pairs"], prm,
                           min_members=5, copy=False)
        torch.cuda.synchronize()
        if i >= 2:
            te += time.perf_counter() - t0
    out = {"ms_per_step": 1e3 * t / args.steps, "e2e_ms_per_step": 1e3 * te / args.steps,
           "hits": len(pts["loc1"]), "numpy_sort_is_scalar": bool(scalar), "numpy_sort_kind": eng.numpy_sort_kind(),
           "rank_on_device": bool(eng.numpy_sort_kind()) and eng.rank_mode == "auto", "blocks": len(blocks), "device_ms": eng.timings()}
    if args.portable_check:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import oracle_blocks
        got = eng.run_resident(prm, min_members=1)
        exp = oracle_blocks(pts, prm)
        same = len(got) == len(exp)
        for b, (pair, d, l1, l2, a, sc) in enumerate(exp if same else []):
            a0, a1 = got.member_off[b], got.member_off[b + 1]
            if not (got.pair[b] == pair and got.dir[b] == d and got.score[b] == sc and got.p_raw[b] == a and
                    (got.loc1[a0:a1] == l1).all() and (got.loc2[a0:a1] == l2).all()):
                same = False
                break
        out["identical_to_oracle_with_this_numpy"] = bool(same)
    eng.close()
    print("PORTABLE " + json.dumps(out), flush=True)


AVX2_NUMPY = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX512_KNL AVX512_KNM"


def run_portable_leg(args, local, rank, world, features=PORTABLE_NUMPY):
    """Parent side: the same workload in a child with numpy's SIMD sorts disabled (see portable_child); with
    `features=AVX2_NUMPY` only the AVX-512 ones, which leaves numpy with its AVX2 argsort."""
    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=features)
    cmd = [sys.executable, os.path.abspath(__file__), "--portable-child", str(local), "--workload", args.workload,
           "--steps", str(args.steps), "--warmup", str(args.warmup), "--seed-shift", str(rank)]
    if rank == 0 and world == 1:
        cmd.append("--portable-check")
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("PORTABLE ")]
        if not line:
            return {"error": (r.stderr or r.stdout)[-300:]}
        return json.loads(line[-1][len("PORTABLE "):])
    except Exception as e:                                  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def strong_leg(args, rank, world, local, barrier, sort_threads):
    """BASELINE configs 3 / 4 and the north_star's target: ONE comparison whose chromosome pairs are LPT-sharded over the N
    ranks (wgdi_b200/shard.py; the reference's analogue is run_colliearity.py:117-141).  Every step is end to end: each
    rank uploads its shard from page-locked host memory, chains, ranks and extracts it, hands its blocks to rank 0 through
    /dev/shm, and rank 0 merges them into the reference's emission order (shard.merge_blocks) — all inside the timed
    region, no collective on the data path.  Returns rank 0's record (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from wgdi_b200 import shard
    from wgdi_b200.engine import Blocks, IclEngine
    FIELDS = shard.BLOCK_FIELDS
    full = make_points(args.strong_workload)                 # same seed on every rank: the same comparison
    prm = params_for(full["cfg"])
    total = len(full["loc1"])
    sizes = np.bincount(full["pair_id"], minlength=full["n_pairs"])
    owner = shard.lpt_assign(shard.pair_cost(sizes), world)
    pts = shard.shard_points(full, rank, world, owner=owner)
    n_local = len(pts["loc1"])
    # host gather without a collective (shard.HostGather): one memory-mapped file in /dev/shm, mapped by every rank once,
    # outside the timed region; per step every rank writes its blocks to their place in the merged arrays and rank 0 ends
    # up holding the whole comparison in the reference's emission order
    shm = f"/dev/shm/icl_strong_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}"
    os.makedirs(shm, exist_ok=True)
    path = os.path.join(shm, "merged.bin")
    P_all = int(full["n_pairs"])
    cap_b, cap_m = total // 5 + 64, 2 * total + 64                   # blocks of >= 5 gene pairs, both directions
    hg = shard.HostGather(path, 0, world, P_all, cap_b, cap_m, create=True) if rank == 0 else None
    barrier()
    if rank != 0:
        hg = shard.HostGather(path, rank, world, P_all, cap_b, cap_m, create=False)
    eng = IclEngine(local, sort_threads=sort_threads)
    eng.fields = "writer"      # every rank brings back and places what the file is written from
    from wgdi_b200.params import compact_points
    host_in = {}
    for k, arr in zip(("pair_id", "loc1", "loc2", "grading"), compact_points(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"])):
        buf = eng._buf("in_" + k, n_local, arr.dtype)
        buf[:] = arr
        host_in[k] = buf
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    phase = {}
    merged = None

    def step(timed):
        nonlocal merged
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        b = eng.chain(host_in["pair_id"], host_in["loc1"], host_in["loc2"], host_in["grading"], pts["n_pairs"], prm,
                      min_members=5, copy=False)
        t1 = time.perf_counter()
        if world > 1:
            merged = hg.gather(b, pts["global_pair"], pts["rows"])
        else:
            merged = b
        t2 = time.perf_counter()
        if timed:
            for k, v in eng.timings().items():
                phase.setdefault(k, []).append(v)
            phase.setdefault("chain_call_ms", []).append((t1 - t0) * 1e3)
            phase.setdefault("gather_merge_ms", []).append((t2 - t1) * 1e3)      # incl. the wait for the slowest rank
        return t2 - t0

    try:
        for _ in range(2):
            step(False)
        T = sum(step(True) for _ in range(args.steps))
    finally:
        eng.close()
    mean = {k: float(np.mean(v)) for k, v in phase.items()}
    keys = sorted(mean)
    vec = torch.tensor([T] + [mean[k] for k in keys] + [float(n_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vec.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = vec.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    else:
        mx = mn = vec
    if rank != 0:
        return None
    mx, mn = mx.tolist(), mn.tolist()
    T = mx[0]
    ph_max = dict(zip(keys, mx[1:-1]))
    dev = {k: v for k, v in ph_max.items() if k not in ("chain_call_ms", "gather_merge_ms")}
    limiting = max(dev, key=dev.get) if dev else None
    return {"workload": workload_label(args.strong_workload) + f", ONE comparison, chromosome pairs LPT-sharded over {world} GPU(s)",
            "scaling": "strong", "value": total * args.steps / T, "unit": "hits/s", "ms_per_step": 1e3 * T / args.steps,
            "graded_hits": total, "chromosome_pairs": int(full["n_pairs"]), "blocks": int(len(merged)),
            "what": "end to end per step: H2D of every rank's shard from pinned host memory, sort, DP, rank, extraction, "
                    "blocks to the host, host gather: every rank writes its blocks to their place in the merged arrays "
                    "(one memory-mapped /dev/shm file, shard.HostGather; no collective on the data path); the step ends "
                    "when rank 0 holds the whole comparison in emission order; max over ranks",
            "hits_per_rank": {"max": int(mx[-1]), "min": int(mn[-1])},
            "phase_ms_max_over_ranks": ph_max, "limiting_device_phase": limiting,
            "lpt_cost": "points + points^2 / 65536 per chromosome pair (shard.pair_cost)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-wall", action="store_true")
    ap.add_argument("--no-reference-sample", action="store_true", help="skip the timed sample of the real reference (oracle/_ref)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg")
    ap.add_argument("--strong-workload", default="C4", help="ONE comparison sharded over all ranks (strong scaling leg)")
    ap.add_argument("--batch", type=int, default=12, help="comparisons in the pipelined-batch leg (0 = skip)")
    ap.add_argument("--portable-child", type=int, default=None, help=argparse.SUPPRESS)
    ap.add_argument("--portable-check", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--seed-shift", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--no-portable", action="store_true")
    ap.add_argument("--portable", action="store_true",
                    help="run the WHOLE benchmark (either arm) with numpy's SIMD sorts disabled (NPY_DISABLE_CPU_FEATURES): "
                         "numpy's scalar argsort, the tie order that is the same on every machine; the engine then ranks on the device")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.portable and os.environ.get("NPY_DISABLE_CPU_FEATURES") != PORTABLE_NUMPY:
        # numpy reads the variable when it is imported, which has happened already: start over with it set
        os.environ["NPY_DISABLE_CPU_FEATURES"] = PORTABLE_NUMPY
        os.execve(sys.executable, [sys.executable] + sys.argv, os.environ)
    if args.portable_child is not None:
        portable_child(args)
        return
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    from wgdi_b200 import shard
    bound_cpus = shard.bind_to_gpu_node(local) if world > 1 else None   # before any page-locked allocation
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from wgdi_b200 import shard
    from wgdi_b200.engine import IclEngine

    # ---- workload: one comparison per rank, pairs pooled and LPT-sharded over ranks --------------------------
    mine = make_points(args.workload, seed_shift=rank)
    prm = params_for(mine["cfg"])
    if world > 1:
        shm = f"/dev/shm/icl_bench_{os.environ.get('MASTER_PORT', '0')}"
        os.makedirs(shm, exist_ok=True)
        np.savez(os.path.join(shm, f"r{rank}.npz"), pair_id=mine["pair_id"], loc1=mine["loc1"], loc2=mine["loc2"],
                 grading=mine["grading"], n_pairs=mine["n_pairs"])
        barrier()
        parts = [np.load(os.path.join(shm, f"r{r}.npz")) for r in range(world)]
        base = np.r_[0, np.cumsum([int(p["n_pairs"]) for p in parts])]
        pooled = dict(pair_id=np.concatenate([p["pair_id"] + base[r] for r, p in enumerate(parts)]).astype(np.int32),
                      loc1=np.concatenate([p["loc1"] for p in parts]), loc2=np.concatenate([p["loc2"] for p in parts]),
                      grading=np.concatenate([p["grading"] for p in parts]), n_pairs=int(base[-1]))
        pts = shard.shard_points(pooled, rank, world)
        total_points = len(pooled["loc1"])
        total_pairs = pooled["n_pairs"]
        del parts, pooled
    else:
        pts = mine
        total_points = len(pts["loc1"])
        total_pairs = pts["n_pairs"]
    n_local = len(pts["loc1"])

    # the host rank (numpy's argsort) is per-rank host work: split the box's cores between the ranks
    sort_threads = int(os.environ.get("ICL_BENCH_SORT_THREADS", 0)) or max(1, (os.cpu_count() or 1) // world)
    eng = IclEngine(local, sort_threads=sort_threads)
    eng.fields = "writer"      # like the drop-in driver: what the collinearity file is written from comes back (icl_extract_fields)
    # inputs in page-locked host memory (the e2e leg copies them every step)
    # (in the narrowest element types that hold them, as the drop-in driver hands them over: params.compact_points)
    from wgdi_b200.params import compact_points
    host_in = {}
    for k, arr in zip(("pair_id", "loc1", "loc2", "grading"), compact_points(pts["pair_id"], pts["loc1"], pts["loc2"], pts["grading"])):
        buf = eng._buf("in_" + k, n_local, arr.dtype)
        buf[:] = arr
        host_in[k] = buf
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    dev_ms = {}

    def step_resident():
        return eng.run_resident(prm, min_members=5, copy=False)         # what the file writer keeps (run_colliearity.py:178)

    def read_phase_clocks():
        # the library's CUDA events of the step that just ended (recorded inside the timed region on its stream); reading
        # them is host work of the benchmark, not of the step: after the step's clock has stopped
        for k, v in eng.timings().items():
            dev_ms.setdefault(k, []).append(v)
        dev_ms.setdefault("host_rank", []).append(eng.host_ms["rank"])

    def step_e2e():
        return eng.chain(host_in["pair_id"], host_in["loc1"], host_in["loc2"], host_in["grading"], pts["n_pairs"], prm, min_members=5, copy=False)

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------
    eng.upload(host_in["pair_id"], host_in["loc1"], host_in["loc2"], host_in["grading"], pts["n_pairs"])
    for _ in range(args.warmup):
        step_resident()
    dev_ms.clear()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # device-side clock: CUDA events on the stream the library launches on (host work between the two device phases
    # leaves that stream idle, so event time and wall time must agree; both are reported)
    estream = torch.cuda.ExternalStream(int(eng.lib.icl_stream(eng.ctx)), device=torch.device("cuda", local))

    def timed(step, after=None):
        wall, evms, out = 0.0, 0.0, None
        for _ in range(args.steps):
            flush_l2()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(estream)
            t0 = time.perf_counter()
            out = step()
            e1.record(estream)
            torch.cuda.synchronize()
            wall += time.perf_counter() - t0
            evms += e0.elapsed_time(e1)
            if after:
                after()
        return wall, evms * 1e-3, out

    t_res, ev_res, blocks = timed(step_resident, read_phase_clocks)
    # ---- e2e: host buffers in, blocks on the host out ----------------------------------------------------------
    for _ in range(2):
        step_e2e()
    res_ms = {k: float(np.mean(v)) for k, v in dev_ms.items()}
    t_e2e, ev_e2e, blocks = timed(step_e2e)
    # ---- the DP kernel alone (export epilogue off), for the roofline note: untimed extra passes -------------------
    dp_only_ms = None
    if getattr(eng, "stream_rank", False):
        eng.stream_rank = False
        step_resident()
        dev_ms.clear()
        for _ in range(3):
            step_resident()
            read_phase_clocks()
        dp_only_ms = float(np.mean(dev_ms["chain"]))
        eng.stream_rank = True
    # ---- the DP kernel's critical path, measured by the kernel itself (ICL_B200_CHAIN_DEBUG: clock64 per work item, the
    # <true, true> instantiation of k_chain_ring): one extra untimed pass.  The kernel ends with its slowest
    # (chromosome pair, direction) item; an item is `rounds` dependent rounds of one warp.
    critical_path = None
    if rank == 0:
        import tempfile
        dbg = tempfile.NamedTemporaryFile(prefix="icl_chain_dbg_", suffix=".txt", delete=False)
        dbg.close()
        os.environ["ICL_B200_CHAIN_DEBUG"] = dbg.name
        try:
            step_resident()
            rows = np.loadtxt(dbg.name, dtype=np.int64, ndmin=2) if os.path.getsize(dbg.name) else np.zeros((0, 8), np.int64)
            rows = rows[rows[:, 4] > 0] if len(rows) else rows
            if len(rows):
                # columns: launch slot, item, points, rows, rounds, cycles, cycles waiting for staged data, records
                h = rows[np.argmax(rows[:, 5])]
                critical_path = {
                    "items": int(len(rows)), "points_heaviest": int(h[2]), "records_heaviest": int(h[7]),
                    "rounds_heaviest": int(h[4]), "cycles_heaviest": int(h[5]),
                    "cycles_per_round": float(h[5]) / float(h[4]), "records_per_round": float(h[7]) / float(h[4]),
                    "wait_share_heaviest": float(h[6]) / float(h[5]),
                    "cycles_median_item": float(np.median(rows[:, 5])),
                    "records_per_round_all_items": float(rows[:, 7].sum()) / float(rows[:, 4].sum()),
                    "what": "per-item clock64 counters of one extra untimed pass (ICL_B200_CHAIN_DEBUG); kernel time = "
                            "cycles_heaviest / SM clock (plus the launch and the clock reads of the debugging build)"}
        except Exception as e:                         # diagnostics only
            critical_path = {"error": f"{type(e).__name__}: {e}"}
        finally:
            os.environ.pop("ICL_B200_CHAIN_DEBUG", None)
            try:
                os.unlink(dbg.name)
            except OSError:
                pass
    # ---- the same steps with numpy itself ranking on the host (what every step did before the device reproduced numpy's
    # argsort, and what still happens where numpy sorts in a way the device does not know): for comparison
    host_ranked = None
    if getattr(eng, "rank_mode", None) == "auto" and hasattr(eng, "numpy_sort_kind") and eng.numpy_sort_kind():
        keep = {k: list(v) for k, v in dev_ms.items()}
        try:
            eng.rank_mode = "host"
            for _ in range(2):
                step_resident()
            t_h, ev_h, _ = timed(step_resident, read_phase_clocks)
            host_rank_ms = float(np.mean(dev_ms["host_rank"][-args.steps:]))
            if world > 1:
                th = torch.tensor([t_h], dtype=torch.float64, device="cuda")
                dist.all_reduce(th, op=dist.ReduceOp.MAX)
                t_h = float(th.item())
            host_ranked = {"value": total_points * args.steps / t_h, "unit": "hits/s", "ms_per_step": 1e3 * t_h / args.steps,
                           "host_rank_ms": host_rank_ms, "sort_threads": sort_threads,
                           "what": "rank by numpy on the host between the two device phases (streamed), same result"}
        except Exception as e:                              # noqa: BLE001 — a comparison leg must not take the run down
            if world > 1:
                raise
            host_ranked = {"error": f"{type(e).__name__}: {e}"}
        finally:
            eng.rank_mode = "auto"
            dev_ms.clear()
            dev_ms.update(keep)
    # ---- NON-REFERENCE option, reported next to the headline and never instead of it: extraction order from a stable
    # device sort (icl_rank_device) instead of numpy's unstable argsort on the host.  Same algorithm, same blocks wherever
    # top scores are distinct, but not the reference's tie order; it shows the device-bound rate of the path.
    device_rank = None
    if hasattr(eng, "run_resident_device_rank"):
        for _ in range(2):
            eng.run_resident_device_rank(prm, min_members=5, copy=False)
        t_dev, ev_dev, out_dev = timed(lambda: eng.run_resident_device_rank(prm, min_members=5, copy=False))
        if world > 1:
            td = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            t_dev = float(td.item())
        device_rank = {"value": total_points * args.steps / t_dev, "unit": "hits/s", "ms_per_step": 1e3 * t_dev / args.steps,
                       "blocks_this_rank": len(out_dev),
                       "what": "NOT bit-exact with the reference: stable device sort instead of numpy's unstable host argsort "
                               "for the extraction order (tied scores in reversed (loc1, loc2) order); inputs resident"}
    # ---- the reference's tie order WITHOUT the host: in a process whose numpy runs its scalar introsort (SIMD sorts
    # disabled — also how the fixtures under tests/golden were generated) the device reproduces numpy's argsort exactly and
    # ranks itself.  One comparison per GPU, every rank's child at the same time.
    avx2_leg = None
    if not args.no_portable and world == 1:
        # third flavour: numpy with its AVX2 argsort (what x86 hosts without AVX-512 run), device-ranked, checked against
        # the oracle ranked by that numpy inside the child
        avx2_leg = run_portable_leg(args, local, rank, world, features=AVX2_NUMPY)
        if "ms_per_step" in avx2_leg:
            avx2_leg["value"] = total_points / (avx2_leg["ms_per_step"] * 1e-3)
            avx2_leg["unit"] = "hits/s"
            avx2_leg["what"] = ("same path in a child process whose numpy has its AVX-512 kernels disabled (AVX2 argsort): "
                                "device-ranked with that flavour, inputs resident")
    portable = None
    if not args.no_portable:
        barrier()
        portable = run_portable_leg(args, local, rank, world)
        if "ms_per_step" in portable:
            tp = torch.tensor([portable["ms_per_step"]], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            portable["ms_per_step_max_over_ranks"] = float(tp.item())
            portable["value"] = total_points / (float(tp.item()) * 1e-3)
            portable["unit"] = "hits/s"
            if "e2e_ms_per_step" in portable:
                te = torch.tensor([portable["e2e_ms_per_step"]], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(te, op=dist.ReduceOp.MAX)
                portable["e2e_value"] = total_points / (float(te.item()) * 1e-3)
            portable["what"] = ("same path in a child process with NPY_DISABLE_CPU_FEATURES (numpy's scalar argsort): the device "
                                "reproduces that argsort exactly and ranks itself; reference-exact for that numpy, inputs resident")
    # ---- batch throughput (BASELINE config 5 shape): independent comparisons through two engines on this GPU, one
    # job's host rank overlapping the next job's upload + device phases (wgdi_b200/batch.py); every job uploads
    batch = None
    if args.batch > 0:
        from wgdi_b200.batch import BatchRunner
        depth = int(os.environ.get("ICL_BENCH_BATCH_DEPTH", "4"))
        runner = BatchRunner(local, depth=depth, sort_threads=sort_threads)
        try:
            jobs = [(dict(host_in, n_pairs=pts["n_pairs"]), prm) for _ in range(args.batch)]   # page-locked inputs, as in the e2e leg
            runner.run(jobs[:2 * depth], min_members=5, consume=lambda i, b: len(b))   # warm-up: buffers, sort pool
            barrier()
            t0 = time.perf_counter()
            outs = runner.run(jobs, min_members=5, consume=lambda i, b: len(b))     # blocks handed over as views, like the file writer takes them
            torch.cuda.synchronize()
            t_batch = time.perf_counter() - t0
            assert all(o == outs[0] for o in outs)
        finally:
            runner.close()
        if world > 1:
            tb = torch.tensor([t_batch], dtype=torch.float64, device="cuda")
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            t_batch = float(tb.item())
        batch = {"value": total_points * args.batch / t_batch, "unit": "hits/s", "comparisons": args.batch * world,
                 "ms_per_comparison": 1e3 * t_batch / args.batch, "engines_per_gpu": depth,
                 "what": f"independent comparisons pipelined through {depth} engines per GPU, each job uploads its points"}
    strong = None
    if not args.no_strong:
        strong = strong_leg(args, rank, world, local, barrier, sort_threads)
    clocks = sampler.stop() if rank == 0 else None
    counters = eng.counters()

    if world > 1:
        t = torch.tensor([t_res, t_e2e, ev_res, ev_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, t_e2e, ev_res, ev_e2e = t.tolist()
        nb = torch.tensor([len(blocks)], dtype=torch.int64, device="cuda")
        dist.all_reduce(nb)
        n_blocks = int(nb.item())
    else:
        n_blocks = len(blocks)

    if rank == 0:
        peak, peak_kind = peaks()
        chain_ms = res_ms["chain"]
        on_device = getattr(eng, "rank_mode", "") == "auto" and hasattr(eng, "numpy_sort_kind") and bool(eng.numpy_sort_kind())
        streamed = bool(getattr(eng, "stream_rank", False)) and counters["chain_kernel"] == "ring" and not on_device
        bytes_pp = DP_BYTES_PER_POINT + (EXPORT_BYTES_PER_POINT if streamed else 0)
        achieved = bytes_pp * n_local / (chain_ms * 1e-3) / 1e9
        # bytes that cross PCIe per e2e step: points in (the four arrays as handed over); blocks out — counted from the arrays
        # that came back (field set "writer": 32 B per block + 8 B per gene pair; "all": 60 + 12) — plus the counts; with
        # the host ranking also the score keys out (2 x 4 B per point: streamed by the DP kernel) and the argsort back in
        from dataclasses import fields as dc_fields
        d2h = sum(int(getattr(blocks, f.name).nbytes) for f in dc_fields(blocks) if getattr(blocks, f.name) is not None) + 64 + \
            (0 if on_device else 8 * n_local)
        h2d = sum(int(v.nbytes) for v in host_in.values()) + (0 if on_device else 8 * n_local)
        line = {
            "metric": METRIC, "value": total_points * args.steps / t_res, "unit": "hits/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
            "ms_per_step_cuda_events": 1e3 * ev_res / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_label(args.workload) + ", one comparison per GPU",
                       "graded_hits": total_points, "chromosome_pairs": total_pairs, "blocks": n_blocks,
                       "mg": list(mine["cfg"].mg), "grading": list(mine["cfg"].grading),
                       "parallelism": f"chromosome pairs LPT-sharded over {world} GPU(s), no collective",
                       "output": f"field set '{eng.fields}' of icl_extract_fields: " + ("pair, score, m, p_raw, member_off, loc1, loc2 "
                                  "(what wgdi/run_colliearity.py:143-155,174-199 writes the file from)" if eng.fields == "writer" else "every array of icl_blocks"),
                       "l2": "flushed between steps (256 MiB memset, untimed); working set also exceeds L2",
                       "tie_order": (f"this process's numpy argsort ({eng.numpy_sort_kind()}) reproduced exactly on the device, checked against numpy at start-up: reference-exact"
                                     if getattr(eng, "rank_mode", "") == "auto" and hasattr(eng, "numpy_sort_kind") and eng.numpy_sort_kind()
                                     else "numpy argsort on host between the two device phases (reference-exact)")},
            "e2e": {"value": total_points * args.steps / t_e2e, "unit": "hits/s", "ms_per_step": 1e3 * t_e2e / args.steps,
                    "ms_per_step_cuda_events": 1e3 * ev_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            # kernels of this library launched inside the timed region, counted at every launch by the library itself
            # (icl_counters; CUB's sort / scan kernels are library code and reported separately)
            "gpu_launches": (counters["launches_dp"] + counters["launches_rank"] + counters["launches_extract"]) * args.steps,
            "cub_calls": counters["cub_calls"] * args.steps,
            "host_rank_fallbacks": int(getattr(eng, "host_rank_fallbacks", 0)),
            "roofline": {"kernel": "k_chain_ring: chain DP, both directions" + (" + streamed score export" if streamed else ""),
                         "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (NCU_TRAFFIC_C2["streamed_export" if streamed else "dp_only"]
                                     if (args.workload == "C2" and world == 1) else None),
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of k_chain_ring on this "
                                           "workload and variant (" + ("profiles/r01_chain_ring_ncu_summary.md" if streamed else
                                                                       "profiles/r02_chain_ring_ncu_summary.md") + ")",
                         "peak_source": peak_kind,
                         "kernel_ms": chain_ms, "algorithmic_bytes": bytes_pp * n_local,
                         "bytes_per_point": bytes_pp,
                         "dp_only": None if dp_only_ms is None else {
                             "kernel_ms": dp_only_ms, "algorithmic_bytes": DP_BYTES_PER_POINT * n_local,
                             "achieved": DP_BYTES_PER_POINT * n_local / (dp_only_ms * 1e-3) / 1e9,
                             "frac": DP_BYTES_PER_POINT * n_local / (dp_only_ms * 1e-3) / 1e9 / peak,
                             "what": "same kernel with the streamed export off (ICL_B200_STREAM=0), 3 extra untimed passes"},
                         "critical_path": critical_path,
                         "note": "order-exact integer DP: bound by the dependent rounds of the heaviest item, not by HBM; see DESIGN.md"},
            "phase_ms": res_ms, "batch": batch, "host_ranked": host_ranked, "device_rank_option": device_rank, "portable_numpy": portable,
            "avx2_numpy": avx2_leg,
            "edges_per_step": counters["edges"], "max_span": counters["max_span"], "chain_kernel": counters["chain_kernel"],
            "clocks": clocks,
            "cpu_binding": ("rank 0 bound to the CPUs next to its GPU: " + ",".join(map(str, bound_cpus))) if bound_cpus else "none",
        }
        blocks_copy = None
        if world == 1 and not args.no_cpu_baseline and not args.no_reference_sample:
            blocks_copy = eng.run_resident(prm, min_members=5, copy=True)           # detached: the engine is closed below
        if strong is not None:
            line["strong"] = strong
        if not args.no_wall and world == 1:
            eng.close()
            line["icl_wall"] = icl_wall_time(args.workload, local)
            if not args.no_strong and args.strong_workload != args.workload:
                try:
                    line["icl_wall_strong_workload"] = icl_wall_time(args.strong_workload, local, bi=False, repeats=1)
                except Exception as e:                      # noqa: BLE001
                    line["icl_wall_strong_workload"] = {"error": f"{type(e).__name__}: {e}"}
        if not args.no_cpu_baseline and world == 1:
            from oracle import c_oracle
            c_oracle.build()
            threads = os.cpu_count() or 1
            cpu_port_pass(mine, prm, threads)
            dt, nb_cpu = cpu_port_pass(mine, prm, threads)
            line["cpu_baseline"] = {"value": len(mine["loc1"]) / dt, "unit": "hits/s", "cores": threads, "kind": "port",
                                    "sample": f"full {args.workload} workload, 1 pass ({len(mine['loc1'])} hits, {nb_cpu} blocks), "
                                              "oracle C port of the reference algorithm, pthread pool over chromosome pairs"}
            if not args.no_reference_sample:
                # the real reference on a bounded stratified sample of this workload's chromosome pairs (about 25 s on all
                # cores), compared block by block with what the CUDA path returned for the same pairs
                try:
                    line["cpu_baseline"]["reference_sample"] = reference_sample(mine, prm, k=min(8, threads), max_points=3000,
                                                                                 gpu_blocks=blocks_copy)
                except Exception as e:                      # noqa: BLE001
                    line["cpu_baseline"]["reference_sample"] = {"error": f"{type(e).__name__}: {e}"}
        try:                                               # the library's block cache over the whole run (icl_cache_stats)
            cs = np.zeros(3, np.int64)
            if eng.lib.icl_cache_stats(cs.ctypes.data, 3) == 3:
                line["block_cache"] = {"served_from_cache": int(cs[0]), "asked_from_driver": int(cs[1]),
                                       "held_mb": round(float(cs[2]) / 2 ** 20, 1),
                                       "what": "device / page-locked blocks of closed contexts reused by later ones in this "
                                               "process (every leg and every -icl run builds its own context)"}
        except Exception as e:                             # noqa: BLE001 — reporting only
            line["block_cache"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
