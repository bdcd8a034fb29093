"""Multi-GPU behind the C ABI in ONE process (brg_index_replicate): the build side is built once and replicated to the other
contexts, the way a DataFusion ExecutionPlan shares its build side across partitions (joins/utils.rs:73-107) and maps
execute(partition) to devices (interval_join.rs:442-550).  With one GPU the replicas live on the same device (the copy path is
the same call); with more, replica i lives on GPU i + 1 and is probed there -- a device that never saw the build columns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(seed=5):
    rng = np.random.default_rng(seed)
    nk, nb, npb = 5, 60_000, 300_007
    bk = rng.integers(0, nk, nb).astype(np.uint32)
    bs = rng.integers(0, 20_000_000, nb).astype(np.int32)
    be = (bs + rng.integers(0, 4000, nb)).astype(np.int32)
    pk = rng.integers(0, nk + 1, npb).astype(np.uint32)
    ps = rng.integers(0, 20_000_000, npb).astype(np.int32)
    pe = (ps + rng.integers(0, 500, npb)).astype(np.int32)
    o = np.lexsort((ps, pk))
    return nk, (bk, bs, be), (pk[o], ps[o], pe[o])


def test_replicas_in_one_process_equal_oracle():
    import torch

    import bio_ranges_b200 as brg
    from bio_ranges_b200.engine import replicate
    from oracle import oracle as O

    n_gpu = torch.cuda.device_count()
    n_rep = 3 if n_gpu == 1 else min(n_gpu - 1, 7)
    devs = [0] * n_rep if n_gpu == 1 else list(range(1, 1 + n_rep))
    nk, b, p = _data()
    pk = p[0].copy()
    pk[pk == nk] = brg.KEY_ABSENT
    p = (pk, p[1], p[2])
    src_ctx = brg.Context(0)
    ix = src_ctx.build(*b, nk)
    ctxs = [brg.Context(d) for d in devs]
    reps = replicate(ix, ctxs)
    assert len(reps) == n_rep and all(r.device_bytes == ix.device_bytes and r.rows == ix.rows for r in reps)
    ix.close()  # replicas are independent of the source
    oi = O.OracleIndex(*b, nk)
    wl, wv, wd = oi.nearest_k1(*p)
    ob, op = O.canonical_pairs(*oi.overlap_pairs(*p))
    for r in reps:
        assert (r.count_overlaps(*p) == oi.count_overlaps(*p)).all()
        assert (r.coverage(*p, strict=True) == oi.coverage(*p, strict=True)).all()
        gl, gv, gd = r.nearest_k1(*p)
        assert (gv == wv).all() and (gl[wv] == wl[wv]).all() and (gd[wv] == wd[wv]).all()
        gb, gp = O.canonical_pairs(*r.overlap_pairs(*p))
        assert len(gb) == len(ob) and (gb == ob).all() and (gp == op).all()
    for r in reps:
        r.close()
    for c in ctxs:
        c.close()
    src_ctx.close()


def test_rows_outlive_their_context():
    """a result released after brg_ctx_destroy (late garbage collection in the host language) must not touch freed memory"""
    import bio_ranges_b200 as brg

    nk, b, p = _data(6)
    ctx = brg.Context(0)
    ix = ctx.build(*b, nk, brg.WANT_PAIRS)
    bi, pi = ix.overlap_pairs(*p, copy=False)  # views of the library's pinned result buffers
    n = len(bi)
    first = int(bi[0]) if n else 0
    ix.close()
    ctx.close()
    assert len(pi) == n and (n == 0 or int(bi[0]) == first)  # still readable: the blocks belong to the result
    del bi, pi  # released now, after the context: goes straight back to the driver
