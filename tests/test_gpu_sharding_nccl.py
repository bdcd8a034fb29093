"""The multi-GPU path on real GPUs (needs >= 2; skipped otherwise; uses every GPU of the box): rank 0 builds, the index slab is replicated with an NCCL
broadcast, every rank probes its contiguous row range, the in-order gather equals the oracle.  The gloo twin of this test
(tests/test_sharding_gloo.py) covers the same plumbing on the CPU box."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _data():
    rng = np.random.default_rng(11)
    nk, nb, npb = 6, 200_000, 1_000_003
    bk = rng.integers(0, nk, nb).astype(np.uint32)
    bs = rng.integers(0, 50_000_000, nb).astype(np.int32)
    be = (bs + rng.integers(0, 5000, nb)).astype(np.int32)
    pk = rng.integers(0, nk, npb).astype(np.uint32)
    ps = rng.integers(0, 50_000_000, npb).astype(np.int32)
    pe = (ps + rng.integers(0, 400, npb)).astype(np.int32)
    o = np.lexsort((ps, pk))  # coordinate-sorted probe table: the tiled kernels on every rank
    return nk, (bk, bs, be), (pk[o], ps[o], pe[o])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import bio_ranges_b200 as brg
    from bio_ranges_b200.sharding import ShardedProbe, gather_rows_in_order, replicate_index

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        dev = f"cuda:{rank}"
        nk, b, p = _data()
        ctx = brg.Context(rank)
        ix = ctx.build(*b, nk) if rank == 0 else None  # only rank 0 has the build side
        ix = replicate_index(ctx, ix, dist, rank, world, dev)
        sp = ShardedProbe(ix, rank, world)
        out = {"count": gather_rows_in_order(sp.count_overlaps(*p), dist, rank, world, dev),
               "coverage": gather_rows_in_order(sp.coverage(*p, strict=True), dist, rank, world, dev)}
        left, valid, d = sp.nearest_k1(*p)
        out["n_left"] = gather_rows_in_order(left, dist, rank, world, dev)
        out["n_valid"] = gather_rows_in_order(valid.astype(np.uint8), dist, rank, world, dev)
        out["n_dist"] = gather_rows_in_order(d, dist, rank, world, dev)
        bi, pi = sp.overlap_pairs(*p)
        out["p_build"] = gather_rows_in_order(bi, dist, rank, world, dev)
        out["p_probe"] = gather_rows_in_order(pi, dist, rank, world, dev)
        if rank == world - 1:  # the last rank reports: it never saw the build side except through the broadcast
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_nccl_replicated_index_and_sharded_probes_equal_oracle():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    from oracle import oracle as O

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = min(torch.cuda.device_count(), 8)  # every GPU of the lease: 2 on a two-GPU box, 8 on the full node
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    r = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    nk, b, p = _data()
    oi = O.OracleIndex(*b, nk)
    assert (r["count"] == oi.count_overlaps(*p)).all()
    assert (r["coverage"] == oi.coverage(*p, strict=True)).all()
    wl, wv, wd = oi.nearest_k1(*p)
    assert (r["n_valid"].astype(bool) == wv).all() and (r["n_left"][wv] == wl[wv]).all() and (r["n_dist"][wv] == wd[wv]).all()
    wb, wp = oi.overlap_pairs(*p)
    assert len(r["p_build"]) == len(wb) and (np.diff(r["p_probe"].astype(np.int64)) >= 0).all()
    gb, gp = O.canonical_pairs(r["p_build"], r["p_probe"])
    ob, op = O.canonical_pairs(wb, wp)
    assert (gb == ob).all() and (gp == op).all()
