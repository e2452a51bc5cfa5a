"""N > 1 path on CPU: two ranks (gloo, world_size 2) shard the chromosome pairs by LPT, run their shard through the
engine (fake backend: the g++ build of the kernel bodies), gather on rank 0 and must reproduce the unsharded result."""
import ctypes
import os
import sys

import numpy as np

from conftest import ROOT


def _worker(rank, world, port, emu_path, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from wgdi_b200 import _lib, shard, synth
    from wgdi_b200.engine import Blocks, ChainParams, IclEngine
    _lib._LIB = _lib._declare(ctypes.CDLL(emu_path))            # fake backend injected by the test
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts = synth.graded_points(synth.make_dataset("tiny"))
    prm = ChainParams(mg1=25, mg2=25, grading=(50, 30, 25))
    sp = shard.shard_points(pts, rank, world)
    eng = IclEngine(0, sort_threads=1)
    b = eng.chain(sp["pair_id"], sp["loc1"], sp["loc2"], sp["grading"], sp["n_pairs"], prm)
    eng.close()
    # the strong-scaling gather of bench.py: every rank places its blocks in the merged arrays of one memory-mapped file
    # (shard.HostGather; its barriers are counters in that file, not collectives), twice to exercise the barrier epochs
    P = pts["n_pairs"]
    path = out_path + ".gather"
    hg = shard.HostGather(path, 0, world, P, 4000, 40000, create=True, timeout_s=120) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        hg = shard.HostGather(path, rank, world, P, 4000, 40000, create=False, timeout_s=120)
    local = Blocks(*[getattr(b, f).copy() for f in shard.BLOCK_FIELDS])
    for _ in range(2):
        placed = hg.gather(local, sp["global_pair"], sp["rows"])
    b.pair = sp["global_pair"][b.pair]
    b.member_idx = sp["rows"][b.member_idx].astype(np.int32)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(b, gathered, dst=0)                       # host-side gather: no data-path collective on GPUs
    if rank == 0:
        merged = shard.merge_blocks(gathered)
        for f in shard.BLOCK_FIELDS:
            assert (getattr(placed, f) == getattr(merged, f)).all(), f     # parallel placement == sequential merge
        np.savez(out_path, pair=merged.pair, dir=merged.dir, score=merged.score, m=merged.m, p_raw=merged.p_raw,
                 member_off=merged.member_off, loc1=merged.loc1, loc2=merged.loc2, member_idx=merged.member_idx)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_equals_whole(tmp_path):
    import torch.multiprocessing as mp
    from conftest import oracle_blocks
    from wgdi_b200 import build, synth
    from wgdi_b200.engine import ChainParams
    emu = build.build_emu(os.path.join(ROOT, "tests", "_emu"))
    out = str(tmp_path / "merged.npz")
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, emu, out), nprocs=2, join=True)
    z = np.load(out)
    pts = synth.graded_points(synth.make_dataset("tiny"))
    exp = oracle_blocks(pts, ChainParams(mg1=25, mg2=25, grading=(50, 30, 25)))
    assert len(exp) == len(z["pair"])
    for b, (pair, d, l1, l2, a, sc) in enumerate(exp):
        a0, a1 = z["member_off"][b], z["member_off"][b + 1]
        assert z["pair"][b] == pair and z["dir"][b] == d and z["score"][b] == sc and z["p_raw"][b] == a
        assert (z["loc1"][a0:a1] == l1).all() and (z["loc2"][a0:a1] == l2).all()
        assert (pts["loc1"][z["member_idx"][a0:a1]] == l1).all()     # member_idx points into the unsharded upload
