"""Sharding of chromosome pairs over GPUs.

The reference distributes chromosome pairs over worker processes in static contiguous chunks
(`n = ceil(len(total) / process)`, wgdi/run_colliearity.py:123-136), which balances badly because the
heavy pairs cluster.  Pairs are independent units of work, so here they are assigned by
longest-processing-time (LPT) on their point count; every rank uploads only its own pairs, there is no
data-path collective, and the per-rank block lists are merged on the host in pair order.
"""
from __future__ import annotations

import heapq

import numpy as np


def bind_to_gpu_node(device: int) -> list[int] | None:
    """One process per GPU: keep this process (and with it the page-locked buffers it allocates from now on) on the CPUs
    next to GPU `device` — /sys/bus/pci/devices/<bus id>/local_cpulist, cut down to the CPUs the process may use.  With
    eight ranks every step moves the points up and the blocks down at the same moment; buffers on the far socket cross the
    inter-socket link twice.  Returns the CPU list it bound to, or None when the topology is not visible (a VM without
    NUMA information, no NVML): nothing is changed then.  Switch off with ICL_B200_BIND=0."""
    import os
    if os.environ.get("ICL_B200_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[device]) if visible and all(v.strip().isdigit() for v in visible.split(",")) else device
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        with open(f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:                                      # noqa: BLE001 — a hint, never a reason to fail
        return None


def lpt_assign(sizes, n_bins: int) -> np.ndarray:
    """owner bin of every pair: pairs in descending size go to the currently least-loaded bin."""
    sizes = np.asarray(sizes, np.int64)
    owner = np.zeros(len(sizes), np.int32)
    heap = [(0, b) for b in range(n_bins)]
    heapq.heapify(heap)
    for p in np.argsort(-sizes, kind="stable"):
        load, b = heapq.heappop(heap)
        owner[p] = b
        heapq.heappush(heap, (load + int(sizes[p]), b))
    return owner


def pair_cost(sizes) -> np.ndarray:
    """Cost of a chromosome pair for the assignment.  Most device phases are linear in the pair's points, but a long pair
    also holds the longest dependent work of the step (the DP sweep of its heaviest window, the top-level partitions of
    its rank segment, its largest predecessor tree), which grows faster than the point count: a quadratic term that equals
    the linear one at 65 536 points keeps two very long pairs off the same GPU when a few pairs dominate."""
    n = np.asarray(sizes, np.float64)
    return np.round(n + n * n / 65536.0).astype(np.int64)


def shard_points(points: dict, rank: int, world: int, owner=None):
    """The sub-problem of `rank`: its pairs renumbered densely, plus the map back to global pair ids."""
    if owner is None:
        sizes = np.bincount(points["pair_id"], minlength=points["n_pairs"])
        owner = lpt_assign(pair_cost(sizes), world)
    mine = np.flatnonzero(owner == rank).astype(np.int32)
    local_of = np.full(points["n_pairs"], -1, np.int32)
    local_of[mine] = np.arange(len(mine), dtype=np.int32)
    sel = np.flatnonzero(owner[points["pair_id"]] == rank)
    return dict(pair_id=local_of[points["pair_id"][sel]], loc1=points["loc1"][sel], loc2=points["loc2"][sel],
                grading=points["grading"][sel], n_pairs=len(mine), global_pair=mine, rows=sel)


def merge_blocks(parts):
    """Merge per-rank `Blocks` (with `.pair` already mapped to global pair ids) into emission order:
    pair ascending; inside a pair the rank-local order (direction, acceptance) is already right."""
    from .engine import Blocks
    parts = [p for p in parts if p is not None]
    pair = np.concatenate([p.pair for p in parts])
    order = np.argsort(pair, kind="stable")
    m = np.concatenate([p.m for p in parts])
    starts = np.concatenate([p.member_off[:-1] + off for p, off in
                             zip(parts, np.cumsum([0] + [len(p.loc1) for p in parts[:-1]]))]) if len(pair) else np.zeros(0, np.int64)
    loc1 = np.concatenate([p.loc1 for p in parts])
    loc2 = np.concatenate([p.loc2 for p in parts])
    midx = np.concatenate([p.member_idx for p in parts])
    mo = m[order].astype(np.int64)
    new_off = np.r_[0, np.cumsum(mo)]
    # member positions of the blocks in their new order: start of the block's run + offset inside it
    gather = (np.repeat(starts[order], mo) + (np.arange(int(new_off[-1]), dtype=np.int64) - np.repeat(new_off[:-1], mo))
              if len(order) else np.zeros(0, np.int64))

    def cat(name):
        return np.concatenate([getattr(p, name) for p in parts])[order]
    return Blocks(pair[order], cat("dir"), cat("score"), m[order], cat("L"), cat("N"), cat("N1"), cat("L1"), cat("L2"),
                  cat("p_raw"), new_off, loc1[gather], loc2[gather], midx[gather])


BLOCK_FIELDS = ("pair", "dir", "score", "m", "L", "N", "N1", "L1", "L2", "p_raw", "member_off", "loc1", "loc2", "member_idx")
MEMBER_FIELDS = ("loc1", "loc2", "member_idx")


def pair_counts(blocks, global_pair, n_pairs: int):
    """(blocks, members) per GLOBAL chromosome pair of one rank's block list."""
    gp = np.asarray(global_pair)[blocks.pair]
    return (np.bincount(gp, minlength=n_pairs).astype(np.int64),
            np.bincount(gp, weights=blocks.m, minlength=n_pairs).astype(np.int64))


def place_blocks(blocks, global_pair, rows, block_base, member_base, out: dict):
    """Write one rank's blocks straight to their place in the merged (emission-order) arrays `out` — the parallel form of
    `merge_blocks`: every chromosome pair is owned by one rank and a rank's blocks are pair-major already, so a pair's run
    of blocks is contiguous on both sides; `block_base` / `member_base` = exclusive sums over the per-pair counts of ALL
    ranks (`pair_counts`).  `rows` maps the rank's upload indices back to the whole comparison's."""
    nb = len(blocks)
    if nb == 0:
        return
    from . import _host
    if _host.available() and out["pair"].dtype == np.int32 and out["member_off"].dtype == np.int64:
        # the same placement by libicl_host.so (icl_host_place_blocks): one pass of memcpy per block instead of numpy's
        # index arrays (half of C4's blocks: 3 ms in numpy, 0.1 ms native)
        bf = [(np.ascontiguousarray(getattr(blocks, f)), out[f]) for f in BLOCK_FIELDS
              if f not in MEMBER_FIELDS and f not in ("pair", "member_off") and getattr(blocks, f) is not None]
        mf = [(np.ascontiguousarray(getattr(blocks, f)), out[f]) for f in ("loc1", "loc2")]
        _host.place_blocks(blocks.pair, global_pair, block_base, member_base, blocks.member_off, bf, mf, out["pair"],
                           out["member_off"], blocks.member_idx, rows if blocks.member_idx is not None else None,
                           out["member_idx"] if blocks.member_idx is not None else None)
        return
    gp = np.asarray(global_pair)[blocks.pair]
    new = np.r_[True, gp[1:] != gp[:-1]]
    first = np.flatnonzero(new)[np.cumsum(new) - 1]            # first block of the pair's run, per block
    dest = block_base[gp] + (np.arange(nb) - first)
    mo = blocks.member_off
    mdest0 = member_base[gp] + (mo[:-1] - mo[first])           # first member of every block in the merged arrays
    for f in BLOCK_FIELDS:
        if f in MEMBER_FIELDS or f == "member_off" or getattr(blocks, f) is None:   # (None: not in the engine's field set)
            continue
        out[f][dest] = gp if f == "pair" else getattr(blocks, f)
    out["member_off"][dest] = mdest0
    shift = np.repeat(mdest0 - mo[:-1], blocks.m)              # merged position - local position, per member
    mdest = shift + np.arange(len(shift))
    out["loc1"][mdest] = blocks.loc1
    out["loc2"][mdest] = blocks.loc2
    if blocks.member_idx is not None:
        out["member_idx"][mdest] = np.asarray(rows)[blocks.member_idx]


class HostGather:
    """Gather of the per-rank block lists of ONE comparison on the host, between processes of one node, without a
    collective: one memory-mapped file (put it in /dev/shm) holds the merged block arrays, the per-rank per-pair counts and
    a barrier word per rank.  Per gather: every rank publishes how many blocks / members each of its chromosome pairs has,
    the ranks meet (`barrier`), every rank derives where each pair's run starts in the merged arrays and writes its own
    blocks there (`place_blocks` — the merge is done by all ranks in parallel), the ranks meet again and rank 0 reads the
    whole comparison in the reference's emission order as views of the file.  The reference's analogue is the parent
    process concatenating the workers' texts in chunk order (wgdi/run_colliearity.py:133-141)."""

    DT = dict(pair=np.int32, dir=np.int32, score=np.int64, m=np.int32, L=np.int32, N=np.int32, N1=np.int64, L1=np.int32,
              L2=np.int32, p_raw=np.float64, member_off=np.int64, loc1=np.int32, loc2=np.int32, member_idx=np.int32)

    def __init__(self, path: str, rank: int, world: int, n_pairs: int, cap_blocks: int, cap_members: int, create: bool,
                 timeout_s: float = 60.0):
        """`create`: this rank makes the (zero-filled) file — call it on one rank, then let the others map it (any
        start-up barrier of the launcher in between)."""
        self.rank, self.world, self.n_pairs, self.timeout_s = int(rank), int(world), int(n_pairs), float(timeout_s)
        size, offs = 0, {}
        for f in ("epoch", "counts") + BLOCK_FIELDS:
            cnt = (8 * world if f == "epoch" else 2 * world * n_pairs if f == "counts" else
                   cap_members if f in MEMBER_FIELDS else cap_blocks + 1)
            dt = np.int64 if f in ("epoch", "counts") else self.DT[f]
            offs[f] = (size, cnt, dt)
            size += (cnt * np.dtype(dt).itemsize + 63) & ~63
        self.cap_blocks, self.cap_members = int(cap_blocks), int(cap_members)
        self.mm = np.memmap(path, dtype=np.uint8, mode="w+" if create else "r+", shape=(size,))
        arrs = {f: np.frombuffer(self.mm, dt, cnt, off) for f, (off, cnt, dt) in offs.items()}
        self.counts = arrs["counts"].reshape(world, 2, n_pairs)
        self.epoch = arrs["epoch"].reshape(world, 8)             # one cache line per rank
        self.out = {f: arrs[f] for f in BLOCK_FIELDS}
        self._epoch = 0

    def barrier(self):
        """The ranks meet in shared memory: a counter per rank (x86: stores are seen in program order, so whatever a rank
        wrote before raising its counter is visible to a rank that has seen the counter)."""
        import time
        self._epoch += 1
        self.epoch[self.rank, 0] = self._epoch
        deadline = time.perf_counter() + self.timeout_s
        while int(self.epoch[:, 0].min()) < self._epoch:
            if time.perf_counter() > deadline:
                raise RuntimeError(f"HostGather: a rank did not reach the barrier within {self.timeout_s} s")

    def gather(self, blocks, global_pair, rows):
        """Collective over the ranks of the file.  blocks: this rank's `Blocks` (pair ids local to its shard);
        global_pair / rows: the shard's maps back to the whole comparison (`shard_points`).  Returns the merged `Blocks`
        (views of the file, valid until the next gather) on rank 0, None elsewhere."""
        from .engine import Blocks
        cb, cm = pair_counts(blocks, global_pair, self.n_pairs)
        self.counts[self.rank, 0], self.counts[self.rank, 1] = cb, cm
        self.barrier()                                           # every rank's counts are in
        bb = np.r_[0, np.cumsum(self.counts[:, 0].sum(0))]
        mb = np.r_[0, np.cumsum(self.counts[:, 1].sum(0))]
        nb, nm = int(bb[-1]), int(mb[-1])
        if nb > self.cap_blocks or nm > self.cap_members:
            raise ValueError("HostGather: the merged block list does not fit the capacity the file was created with")
        place_blocks(blocks, global_pair, rows, bb, mb, self.out)
        if self.rank == 0:
            self.out["member_off"][nb] = nm
        self.barrier()                                           # every rank's blocks are in place
        if self.rank != 0:
            return None
        return Blocks(*[None if getattr(blocks, f) is None else
                        self.out[f][:(nm if f in MEMBER_FIELDS else nb + 1 if f == "member_off" else nb)] for f in BLOCK_FIELDS])


def chain_on_devices(points: dict, n_pairs: int, params, devices, min_members: int = 1):
    """One comparison over several GPUs of this host from ONE process: chromosome pairs LPT-sharded over `devices`
    (`shard_points`), one engine and one host thread per GPU (the C calls release the GIL; the sort pool serves one
    shard at a time, so a shard's device phases overlap the other shards' host rank), blocks merged in pair order.
    Same result as a single engine (tests).  No collective: results meet on the host."""
    import threading

    from .engine import IclEngine
    pts = dict(points, n_pairs=n_pairs)
    world = len(devices)
    parts, errors = [None] * world, []

    def work(r):
        try:
            sp = shard_points(pts, r, world)
            eng = IclEngine(int(devices[r]))
            try:
                b = eng.chain(sp["pair_id"], sp["loc1"], sp["loc2"], sp["grading"], sp["n_pairs"], params,
                              min_members=min_members, copy=True)
            finally:
                eng.close()
            b.pair = sp["global_pair"][b.pair]
            b.member_idx = sp["rows"][b.member_idx].astype(np.int32)
            parts[r] = b
        except Exception as e:                             # noqa: BLE001 — re-raised in the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return merge_blocks(parts)
