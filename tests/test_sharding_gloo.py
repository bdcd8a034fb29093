"""Multi-GPU plumbing on CPU: world_size-2 `gloo` processes run the sharded probe path (replicate the index from rank 0,
contiguous probe-row shards, in-order gather) with the ORACLE standing in for the device index, and the gathered result
must equal the single-process result row for row (mirrors CollectLeft, interval_join.rs:396-399,466-480)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_rows_in_order():
    from bio_ranges_b200.sharding import shard_bounds, shard_range

    for n in (0, 1, 7, 8, 1000, 125_000_001):
        for w in (1, 2, 3, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
            assert shard_range(n, w - 1, w) == b[-1]
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


class _CpuIndex:
    """device_bytes / export() / probe protocol of engine.Index over the oracle.  Like the real slab, the byte tensor starts
    with its own header (row count, key count) followed by the build columns.  Only the replication + sharding logic is under
    test here."""

    def __init__(self, key, start, end, n_keys):
        from oracle import oracle as O

        self.cols = (np.ascontiguousarray(key, np.uint32), np.ascontiguousarray(start, np.int32),
                     np.ascontiguousarray(end, np.int32))
        self.n_keys = n_keys
        self.o = O.OracleIndex(*self.cols, n_keys)

    @property
    def device_bytes(self):
        return 16 + 12 * len(self.cols[0])

    def export(self):
        import torch

        n = len(self.cols[0])
        header = np.array([n, self.n_keys], np.uint64).tobytes()
        slab = torch.from_numpy(np.concatenate([np.frombuffer(header, np.uint8)] + [c.view(np.uint8) for c in self.cols]).copy())
        return header, slab

    def __getattr__(self, name):
        return getattr(self.o, name)


class _CpuCtx:
    def attach_tensor(self, slab):
        raw = slab.numpy()
        n, n_keys = (int(x) for x in raw[:16].view(np.uint64))
        assert len(raw) == 16 + 12 * n
        body = raw[16:]
        return _CpuIndex(body[: 4 * n].view(np.uint32), body[4 * n: 8 * n].view(np.int32), body[8 * n: 12 * n].view(np.int32), n_keys)


def _data():
    rng = np.random.default_rng(7)
    nk, nb, npb = 5, 3000, 10001
    bk = rng.integers(0, nk, nb).astype(np.uint32)
    bs = rng.integers(0, 200000, nb).astype(np.int32)
    be = (bs + rng.integers(0, 900, nb)).astype(np.int32)
    pk = rng.integers(0, nk, npb).astype(np.uint32)
    ps = rng.integers(0, 200000, npb).astype(np.int32)
    pe = (ps + rng.integers(0, 300, npb)).astype(np.int32)
    return nk, (bk, bs, be), (pk, ps, pe)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from bio_ranges_b200.sharding import ShardedProbe, gather_rows_in_order, replicate_index

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nk, b, p = _data()
        ix = _CpuIndex(*b, nk) if rank == 0 else None  # only rank 0 has the build side
        ix = replicate_index(_CpuCtx(), ix, dist, rank, world, "cpu")
        sp = ShardedProbe(ix, rank, world)
        out = {
            "count": gather_rows_in_order(sp.count_overlaps(*p), dist, rank, world),
            "coverage": gather_rows_in_order(sp.coverage(*p, strict=True), dist, rank, world),
        }
        left, valid, d = sp.nearest_k1(*p)
        out["n_left"] = gather_rows_in_order(left, dist, rank, world)
        out["n_valid"] = gather_rows_in_order(valid.astype(np.uint8), dist, rank, world)
        out["n_dist"] = gather_rows_in_order(d, dist, rank, world)
        bi, pi = sp.overlap_pairs(*p)
        out["p_build"] = gather_rows_in_order(bi, dist, rank, world)
        out["p_probe"] = gather_rows_in_order(pi, dist, rank, world)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world,how", [(2, "broadcast"), (4, "broadcast"), (4, "spread")])  # (spread: scatter + all-gather, sharding._spread)
def test_sharded_probe_equals_single_process(world, how, monkeypatch):
    import torch.multiprocessing as mp

    from oracle import oracle as O

    monkeypatch.setenv("BRG_REPLICATE", how)  # (inherited by the spawned ranks)
    port = _free_port()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    nk, b, p = _data()
    oi = O.OracleIndex(*b, nk)
    want_left, want_valid, want_d = oi.nearest_k1(*p)
    wb, wp = oi.overlap_pairs(*p)
    for rank in range(world):  # every rank holds the full, in-order result
        r = results[rank]
        assert (r["count"] == oi.count_overlaps(*p)).all()
        assert (r["coverage"] == oi.coverage(*p, strict=True)).all()
        assert (r["n_left"] == want_left).all() and (r["n_valid"] == want_valid).all() and (r["n_dist"] == want_d).all()
        # pairs: shard-order concatenation is probe-major like the single-process output
        assert len(r["p_build"]) == len(wb)
        assert (np.diff(r["p_probe"].astype(np.int64)) >= 0).all()
        gb, gp = O.canonical_pairs(r["p_build"], r["p_probe"])
        ob, op = O.canonical_pairs(wb, wp)
        assert (gb == ob).all() and (gp == op).all()
