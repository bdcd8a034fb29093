"""BASELINE.json configs[1], [2], [3] (and the configs[4] per-GPU shard) at FULL size, with bench.py's own generators:
every row of the CUDA result against the oracle (pairs: per-probe counts + the order-independent checksum of the pair set,
the scheme of the reference's large-table test, tests/integration_test.rs:4289-4404).  Device-resident inputs, the *_dev
entry points of the C ABI -- the calls bench.py times."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def env():
    import torch

    import bench
    import bio_ranges_b200 as brg

    ctx = brg.Context(0)
    yield torch, bench, brg, ctx
    ctx.close()


def _np(t):
    return t.cpu().numpy()


def _setup(env, name, order):
    torch, bench, brg, ctx = env
    b, p = bench.make_workload(torch, name, "cuda:0", 0, 1, order, 1.0)
    want = {"count": brg.WANT_COUNT, "pairs": brg.WANT_PAIRS, "coverage": brg.WANT_COVERAGE, "nearest": brg.WANT_NEAREST}[name]
    ix = ctx.build_dev(*b, len(bench.HG38), want)
    oi = O.OracleIndex(_np(b[0]).astype(np.uint32), _np(b[1]), _np(b[2]), len(bench.HG38))
    pnp = (_np(p[0]).astype(np.uint32), _np(p[1]), _np(p[2]))
    return torch, ix, oi, p, pnp


@pytest.mark.timeout(600)
@pytest.mark.parametrize("order", ["sorted", "shuffled"])
def test_config1_pairs_10m_x_10m(env, order):
    torch, ix, oi, (pk, ps, pe), pnp = _setup(env, "pairs", order)
    n = ps.numel()
    offsets = torch.empty((n + 255) // 256 + 1, dtype=torch.int64, device="cuda")
    ranges = torch.empty(n, dtype=torch.int64, device="cuda")
    ix.pairs_chunk_offsets_dev(pk, ps, pe, offsets, ranges)
    torch.cuda.synchronize()
    total = int(offsets[-1].item())
    bi = torch.empty(total, dtype=torch.int32, device="cuda")
    pi = torch.empty(total, dtype=torch.int32, device="cuda")
    ix.pairs_fill_dev(ps, offsets, ranges, bi, pi)
    torch.cuda.synchronize()
    want, cnt = oi.overlap_checksum(*pnp, threads=O.max_threads(), counts=True)
    assert 4.0 < total / n < 6.5  # "~5 pairs/probe"
    assert total == want[3]
    gb, gp = _np(bi).view(np.uint32), _np(pi).view(np.uint32)
    assert (np.diff(gp.astype(np.int64)) >= 0).all()  # probe-major
    assert (np.bincount(gp, minlength=n).astype(np.uint64) == cnt).all()
    assert O.pair_checksum(gb, gp) == want
    # the host-buffer call returns the same set
    hb, hp = ix.overlap_pairs(*pnp, copy=False)
    assert len(hb) == total and O.pair_checksum(hb, hp) == want
    del hb, hp
    # the join in one launch (brg_pairs_join_dev): any-order output slots, then ordered; and through the host-buffer call
    import bio_ranges_b200 as brg

    tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    for ordered in (False, True):
        bi.fill_(-1)
        pi.fill_(-1)
        ix.pairs_join_dev(pk, ps, pe, bi, pi, tot, ordered=ordered)
        torch.cuda.synchronize()
        assert int(tot.item()) == total
        gb, gp = _np(bi).view(np.uint32), _np(pi).view(np.uint32)
        if ordered:
            assert (np.diff(gp.astype(np.int64)) >= 0).all()
        assert (np.bincount(gp, minlength=n).astype(np.uint64) == cnt).all()
        assert O.pair_checksum(gb, gp) == want
    hb, hp = ix.overlap_pairs(*pnp, mode=brg.PAIRS_INNER_ANY_ORDER, copy=False)
    assert len(hb) == total and O.pair_checksum(hb, hp) == want
    ix.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("order", ["sorted", "shuffled"])
def test_config2_coverage_100m_x_1m(env, order):
    torch, ix, oi, (pk, ps, pe), pnp = _setup(env, "coverage", order)
    out = torch.empty(ps.numel(), dtype=torch.int64, device="cuda")
    ix.coverage_dev(pk, ps, pe, out)
    torch.cuda.synchronize()
    want = oi.coverage(*pnp, threads=O.max_threads())
    got = _np(out)
    assert (got == want).all()
    assert 0.02 < float((want > 0).mean()) < 0.9  # the workload exercises both covered and uncovered reads
    ix.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("order", ["sorted", "shuffled"])
def test_config3_nearest_50m_x_5m(env, order):
    torch, ix, oi, (pk, ps, pe), pnp = _setup(env, "nearest", order)
    n = ps.numel()
    left = torch.empty(n, dtype=torch.int32, device="cuda")
    bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    dist = torch.empty(n, dtype=torch.int64, device="cuda")
    ix.nearest_k1_dev(pk, ps, pe, left, bits, dist)
    torch.cuda.synchronize()
    ol, ov, od = oi.nearest_k1(*pnp, threads=O.max_threads())
    valid = np.unpackbits(_np(bits).view(np.uint8), bitorder="little")[:n].astype(bool)
    assert (valid == ov).all()
    assert (_np(left).view(np.uint32)[ov] == ol[ov]).all() and (_np(dist)[ov] == od[ov]).all()
    assert float((od[ov] > 0).mean()) > 0.5  # "many non-overlapping": the neighbour search is what runs
    ix.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("order", ["sorted", "shuffled"])
def test_config4_shard_count_125m_x_10m(env, order):
    torch, ix, oi, (pk, ps, pe), pnp = _setup(env, "count", order)
    out = torch.empty(ps.numel(), dtype=torch.int64, device="cuda")
    ix.count_overlaps_dev(pk, ps, pe, out)
    torch.cuda.synchronize()
    assert (_np(out) == oi.count_overlaps(*pnp, threads=O.max_threads())).all()
    ix.close()
