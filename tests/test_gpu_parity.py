"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C ABI (host-buffer entry points, as the
Rust host would call them) and is compared bit-exactly with the CPU oracle and with the reference's golden vectors."""
import zlib

import numpy as np
import pytest

import reference_cases as rc
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import bio_ranges_b200 as brg

    c = brg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def make_index(ctx):
    def _mk(key, start, end, n_keys):
        return ctx.build(key, np.asarray(start, np.int32), np.asarray(end, np.int32), n_keys)

    return _mk


@pytest.mark.parametrize("case", rc.ALL_CASES, ids=lambda c: c.__name__)
def test_gpu_reference_case(case, make_index):
    """the reference's own unit literals / CSV tables / golden parquet, on the CUDA path"""
    case(make_index)


def _rand(rng, n, n_keys, lo, hi, maxlen, absent_frac=0.0):
    k = rng.integers(0, n_keys, n).astype(np.uint32)
    s = rng.integers(lo, hi, n).astype(np.int64)
    e = s + rng.integers(0, maxlen, n)
    if absent_frac:
        k[rng.random(n) < absent_frac] = O.ABSENT
    return k, s.astype(np.int32), np.clip(e, -2**31, 2**31 - 1).astype(np.int32)


def _compare_all(ctx, bk, bs, be, pk, ps, pe, nk, ks=(2, 5)):
    gi = ctx.build(bk, bs, be, nk)
    oi = O.OracleIndex(bk, bs, be, nk)
    for strict in (False, True):
        assert (gi.count_overlaps(pk, ps, pe, strict=strict) == oi.count_overlaps(pk, ps, pe, strict=strict)).all()
        assert (gi.coverage(pk, ps, pe, strict=strict) == oi.coverage(pk, ps, pe, strict=strict)).all()
        for inc in (True, False):
            gl, gv, gd = gi.nearest_k1(pk, ps, pe, include_overlaps=inc, strict=strict)
            ol, ov, od = oi.nearest_k1(pk, ps, pe, include_overlaps=inc, strict=strict)
            assert (gv == ov).all()
            assert (gl == ol).all() and (gd == od).all()
            for k in ks:
                g = gi.nearest_k(pk, ps, pe, k, include_overlaps=inc, strict=strict)
                o = oi.nearest_k(pk, ps, pe, k, include_overlaps=inc, strict=strict)
                assert len(g[0]) == len(o[0])
                for a, b in zip(g, o):
                    assert (a == b).all()  # same rows in the same order (probe-major, reference emission order)
    gb, gp = O.canonical_pairs(*gi.overlap_pairs(pk, ps, pe))
    ob, op = O.canonical_pairs(*oi.overlap_pairs(pk, ps, pe))
    assert len(gb) == len(ob) and (gb == ob).all() and (gp == op).all()
    cnt = oi.overlap_counts(pk, ps, pe)
    import bio_ranges_b200 as brg

    semi = gi.overlap_pairs(pk, ps, pe, mode=brg.PAIRS_SEMI)
    anti = gi.overlap_pairs(pk, ps, pe, mode=brg.PAIRS_ANTI)
    assert (semi == np.nonzero(cnt > 0)[0]).all() and (anti == np.nonzero(cnt == 0)[0]).all()
    gi.close()


@pytest.mark.parametrize("seed", range(8))
def test_gpu_vs_oracle_random(ctx, seed):
    rng = np.random.default_rng(seed)
    nk = int(rng.choice([1, 3, 24, 300]))
    nb = int(rng.choice([1, 50, 5000, 60000]))
    maxlen = int(rng.choice([1, 40, 5000, 200000]))
    span = int(rng.choice([1000, 10**6, 2 * 10**9]))
    bk, bs, be = _rand(rng, nb, nk, -span if seed % 2 else 0, span, maxlen)
    pk, ps, pe = _rand(rng, 20000, nk, -span if seed % 2 else 0, span, int(rng.choice([1, 200, 3000])), absent_frac=0.05)
    _compare_all(ctx, bk, bs, be, pk, ps, pe, nk)


def _sorted_probes(rng, n, nk, span, len_lo, len_hi, absent_runs=False, block=None):
    """coordinate-sorted probe table (by key, then start); ends are NOT sorted unless len_lo == len_hi"""
    k = rng.integers(0, nk, n).astype(np.uint32)
    s = rng.integers(0, span, n).astype(np.int64)
    e = s + rng.integers(len_lo, len_hi + 1, n)
    if block:  # sorted inside blocks of `block` rows only (e.g. concatenated sorted partitions)
        o = np.concatenate([b0 + np.lexsort((s[b0:b0 + block], k[b0:b0 + block])) for b0 in range(0, n, block)])
    else:
        o = np.lexsort((s, k))
    k, s, e = k[o], s[o], e[o]
    if absent_runs:  # a few runs of rows whose key is not on the build side
        for b0 in rng.integers(0, max(1, n - 700), 5):
            k[b0:b0 + int(rng.integers(1, 700))] = O.ABSENT
    return k, s.astype(np.int32), np.clip(e, -2**31, 2**31 - 1).astype(np.int32)


SORTED_REGIMES = {
    # name: (n_build, n_probe, n_keys, span, build_len, probe_len_lo, probe_len_hi, kwargs)
    "fixed_len_sparse_build": (3_000, 70_003, 3, 2_000_000, 400, 150, 150, {}),
    "variable_len": (40_000, 50_001, 4, 3_000_000, 900, 1, 1500, {}),
    "build_denser_than_probes": (200_000, 20_000, 2, 1_000_000, 60, 1, 300, {}),
    "build_much_denser": (400_000, 4_099, 1, 400_000, 30, 1, 100, {}),
    "long_probes": (60_000, 30_000, 2, 2_000_000, 200, 1, 60_000, {}),
    "absent_key_runs": (30_000, 40_000, 5, 1_000_000, 500, 10, 800, {"absent_runs": True}),
    "sorted_blocks": (30_000, 48_000, 3, 1_000_000, 500, 10, 800, {"block": 6000}),
    "many_keys_short_runs": (20_000, 30_000, 300, 100_000, 300, 1, 500, {}),
    "tiny": (50, 300, 2, 10_000, 100, 1, 300, {}),
}


@pytest.mark.parametrize("regime", list(SORTED_REGIMES), ids=list(SORTED_REGIMES))
def test_gpu_sorted_probe_regimes(ctx, regime):
    """the TMA-tiled fast paths: start-sorted probe tables in the regimes that select different search strategies
    (inverse search on ends / on starts + forward scan / wide windows / deferred chunks)"""
    nb, npb, nk, span, blen, pl0, pl1, kw = SORTED_REGIMES[regime]
    rng = np.random.default_rng(zlib.crc32(regime.encode()))
    bk, bs, be = _rand(rng, nb, nk, 0, span, blen)
    pk, ps, pe = _sorted_probes(rng, npb, nk, span, pl0, pl1, **kw)
    _compare_all(ctx, bk, bs, be, pk, ps, pe, nk, ks=(3,))


@pytest.mark.parametrize("pile", [1, 31, 32, 33, 96, 126, 127, 128, 129, 200])
def test_gpu_window_piles_in_one_histogram_bin(ctx, pile):
    """inverse search edge: `pile` build elements share one insertion point among a chunk's 256 sorted queries (the histogram
    bins are bytes; a window holds at most 127 elements before the chunk is handed to the per-row kernel), in the first,
    a middle and the last bin, for the start-, end- and prefix-max-keyed searches"""
    rng = np.random.default_rng(pile)
    n_rows = 256 * 6
    for where in (0, 100, 255):
        # sorted probes, fixed length (ends monotone too); the pile sits between row `where - 1` and row `where` of chunk 2
        ps = np.arange(n_rows, dtype=np.int64) * 1000 + 10_000
        pe = ps + 200
        gap_lo = ps[512 + where] - 700 if where else ps[511] + 300
        bs_pile = gap_lo + rng.integers(0, 50, pile)            # starts piled into one gap between consecutive probes
        be_pile = bs_pile + rng.integers(0, 40, pile)
        bs_far = rng.integers(0, int(ps[-1]) + 5000, 300)       # a sparse background
        be_far = bs_far + rng.integers(0, 3000, 300)
        bs = np.concatenate([bs_pile, bs_far]).astype(np.int32)
        be = np.concatenate([be_pile, be_far]).astype(np.int32)
        bk = np.zeros(len(bs), np.uint32)
        pk = np.zeros(n_rows, np.uint32)
        _compare_all(ctx, bk, bs, be, pk, ps.astype(np.int32), pe.astype(np.int32), 1, ks=(2,))


def test_gpu_count_with_inverted_build_rows(ctx):
    """build rows with start > end: `(started - ended_before) as i64` (interval_tree.rs:46-48) goes negative where more ends
    than starts precede the probe; sorted probes (the tiled kernel's transposed 32-bit staging) and shuffled ones (per row)"""
    rng = np.random.default_rng(99)
    nk, nb, npb = 3, 20_000, 60_000
    bk, bs, be = _rand(rng, nb, nk, 0, 1_000_000, 400)
    inv = rng.random(nb) < 0.3
    bs2, be2 = np.where(inv, be + 5000, bs).astype(np.int32), np.where(inv, bs, be).astype(np.int32)  # start > end on 30 % of the rows
    pk, ps, pe = _sorted_probes(rng, npb, nk, 1_000_000, 1, 300)
    gi = ctx.build(bk, bs2, be2, nk, brg_want_count())
    oi = O.OracleIndex(bk, bs2, be2, nk)
    want = oi.count_overlaps(pk, ps, pe)
    assert (want < 0).any() and (want > 0).any()
    assert (gi.count_overlaps(pk, ps, pe) == want).all()
    perm = rng.permutation(npb)
    assert (gi.count_overlaps(pk[perm], ps[perm], pe[perm]) == want[perm]).all()
    gi.close()


def brg_want_count():
    import bio_ranges_b200 as brg

    return brg.WANT_COUNT


def test_gpu_pairs_zero_copy_views(ctx):
    """overlap_pairs(copy=False): numpy views of the library's pinned result buffers, released with the last view"""
    import gc

    rng = np.random.default_rng(3)
    bk, bs, be = _rand(rng, 5000, 2, 0, 200_000, 500)
    pk, ps, pe = _rand(rng, 20_000, 2, 0, 200_000, 300)
    gi = ctx.build(bk, bs, be, 2)
    b0, p0 = gi.overlap_pairs(pk, ps, pe)
    b1, p1 = gi.overlap_pairs(pk, ps, pe, copy=False)
    assert (b0 == b1).all() and (p0 == p1).all() and not b1.flags.owndata
    keep = p1[10:20].copy()
    tail = b1[-5:]  # a slice keeps the buffers alive after the parents are gone
    del b1, p1
    gc.collect()
    assert (tail == b0[-5:]).all() and (keep == p0[10:20]).all()
    del tail
    gc.collect()
    for _ in range(3):  # buffers return to the pool and are handed out again
        b2, p2 = gi.overlap_pairs(pk, ps, pe, copy=False)
        assert (b2 == b0).all() and (p2 == p0).all()
    gi.close()


def test_gpu_key_runs_entry_points(ctx):
    """run-length key column (what the host's contig-cache loop yields) == per-row key column"""
    import bio_ranges_b200 as brg

    rng = np.random.default_rng(21)
    nk = 6
    bk, bs, be = _rand(rng, 30_000, nk, 0, 1_000_000, 700)
    pk, ps, pe = _sorted_probes(rng, 100_003, nk, 1_000_000, 1, 600, absent_runs=True)
    rk, rend = brg._cabi.key_runs(pk)
    assert rend[-1] == len(pk) and len(rk) < 40
    gi = ctx.build(bk, bs, be, nk)
    for strict in (False, True):
        assert (gi.count_overlaps_runs(rk, rend, ps, pe, strict=strict) == gi.count_overlaps(pk, ps, pe, strict=strict)).all()
        assert (gi.coverage_runs(rk, rend, ps, pe, strict=strict) == gi.coverage(pk, ps, pe, strict=strict)).all()
        a = gi.nearest_k1_runs(rk, rend, ps, pe, strict=strict)
        b = gi.nearest_k1(pk, ps, pe, strict=strict)
        assert all((x == y).all() for x, y in zip(a, b))
    oi = O.OracleIndex(bk, bs, be, nk)
    assert (gi.count_overlaps_runs(rk, rend, ps, pe) == oi.count_overlaps(pk, ps, pe)).all()
    with pytest.raises(brg.BioRangesError):
        gi.count_overlaps_runs(rk[:-1], rend[:-1], ps, pe)  # runs do not cover the rows
    gi.close()


@pytest.mark.parametrize("dtypes", [(np.int64, np.int64), (np.uint32, np.uint64), (np.int32, np.int64), (np.uint64, np.int32)],
                         ids=lambda d: f"{np.dtype(d[0]).name}-{np.dtype(d[1]).name}")
def test_gpu_arrow_ingest_native_width_columns(ctx, dtypes):
    """position columns cross PCIe in their Arrow width; PosArray::resolve (array_utils.rs:73-131) runs on
    the device.  Same answers as the i32 entry points; values that do not fit fail with the reference's text and first row."""
    import bio_ranges_b200 as brg

    rng = np.random.default_rng(77)
    nk = 5
    bk, bs, be = _rand(rng, 20_000, nk, 0, 2_000_000, 900)
    pk, ps, pe = _sorted_probes(rng, 70_001, nk, 2_000_000, 1, 400, absent_runs=True)  # odd length: vector body + scalar tail
    rk, rend = brg._cabi.key_runs(pk)
    gi = ctx.build(bk, bs, be, nk)
    s_t, e_t = ps.astype(dtypes[0]), pe.astype(dtypes[1])
    for strict in (False, True):
        assert (gi.probe_arrow("count", rk, rend, s_t, e_t, strict=strict) == gi.count_overlaps(pk, ps, pe, strict=strict)).all()
        assert (gi.probe_arrow("coverage", rk, rend, s_t, e_t, strict=strict) == gi.coverage(pk, ps, pe, strict=strict)).all()
        a = gi.probe_arrow("nearest", rk, rend, s_t, e_t, strict=strict)
        b = gi.nearest_k1(pk, ps, pe, strict=strict)
        assert all((x == y).all() for x, y in zip(a, b))
    # overflow: first offending row of the start column wins over the end column (array_utils.rs:94-98)
    big = 2**31
    wide = [i for i, d in enumerate(dtypes) if np.dtype(d) != np.int32]
    cols = [s_t.copy(), e_t.copy()]
    rows = {0: (60_001, 12_345), 1: (70_000, 999)}  # (a row in the scalar tail / last vector, an earlier row)
    for c in wide:
        for r in rows[c]:
            cols[c][r] = big + c
    first_col = wide[0]
    with pytest.raises(brg.BioRangesError) as ei:
        gi.probe_arrow("count", rk, rend, cols[0], cols[1])
    assert ei.value.status == 4  # BRG_ERR_OVERFLOW
    assert f"coordinate value {big + first_col} at row {min(rows[first_col])} overflows i32 (max 2147483647)" in str(ei.value)
    if np.dtype(dtypes[0]) == np.int64:  # negative values below i32::MIN are caught as well
        s_neg = s_t.copy()
        s_neg[7] = -(2**31) - 1
        with pytest.raises(brg.BioRangesError, match=rf"coordinate value {-(2**31) - 1} at row 7 overflows i32"):
            gi.probe_arrow("coverage", rk, rend, s_neg, e_t)
    gi.close()


def test_gpu_host_buffer_chunk_schedule(monkeypatch):
    """host-buffer entry points stream the probe table through double-buffered staging chunks: with a small chunk size the
    schedule has several full chunks and an odd tail"""
    import bio_ranges_b200 as brg

    monkeypatch.setenv("BRG_CHUNK_ROWS", "262144")
    c2 = brg.Context(0)
    rng = np.random.default_rng(5)
    nk = 4
    bk, bs, be = _rand(rng, 50_000, nk, 0, 5_000_000, 800)
    n = 262144 * 3 + 262144 // 2 + 77
    pk, ps, pe = _sorted_probes(rng, n, nk, 5_000_000, 1, 300, absent_runs=True)
    rk, rend = brg._cabi.key_runs(pk)
    gi = c2.build(bk, bs, be, nk)
    oi = O.OracleIndex(bk, bs, be, nk)
    assert (gi.count_overlaps(pk, ps, pe) == oi.count_overlaps(pk, ps, pe)).all()
    assert (gi.count_overlaps_runs(rk, rend, ps, pe) == oi.count_overlaps(pk, ps, pe)).all()
    assert (gi.coverage_runs(rk, rend, ps, pe) == oi.coverage(pk, ps, pe)).all()
    left, valid, dist = gi.nearest_k1_runs(rk, rend, ps, pe)
    ol, ov, od = oi.nearest_k1(pk, ps, pe)
    assert (valid == ov).all() and (left[valid] == ol[ov]).all() and (dist[valid] == od[ov]).all()
    assert (gi.probe_arrow("count", rk, rend, ps.astype(np.int64), pe.astype(np.uint32)) == oi.count_overlaps(pk, ps, pe)).all()
    gi.close()
    c2.close()


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
@pytest.mark.parametrize("order", ["sorted", "shuffled"])
def test_gpu_host_buffers_pageable_and_pinned(monkeypatch, pinned, order):
    """host-buffer entry points with ordinary heap arrays (moved through the pinned staging ring by the host threads) and with
    page-locked arrays (used in place); a per-row key column is sent as per-chunk runs when the table is sorted and as the
    column itself when a chunk holds more than 1024 runs"""
    import torch

    import bio_ranges_b200 as brg

    monkeypatch.setenv("BRG_CHUNK_ROWS", "131072")
    monkeypatch.setenv("BRG_HOST_THREADS", "4")
    c2 = brg.Context(0)
    rng = np.random.default_rng(9)
    nk = 6
    bk, bs, be = _rand(rng, 80_000, nk, 0, 8_000_000, 900)
    n = 131072 * 4 + 4321
    pk, ps, pe = _sorted_probes(rng, n, nk, 8_000_000, 1, 400, absent_runs=True)
    if order == "shuffled":
        o = rng.permutation(n)
        pk, ps, pe = pk[o], ps[o], pe[o]

    def host(a):
        if not pinned:
            return np.ascontiguousarray(a)
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.int32 if a.dtype == np.uint32 else a.dtype)).pin_memory()
        return t.numpy().view(a.dtype)

    pk_h, ps_h, pe_h = host(pk), host(ps), host(pe)
    gi = c2.build(bk, bs, be, nk)
    oi = O.OracleIndex(bk, bs, be, nk)
    out = host(np.zeros(n, np.int64))
    got = gi.count_overlaps(pk_h, ps_h, pe_h, out=out)
    assert got is out and (out == oi.count_overlaps(pk, ps, pe)).all()
    assert (gi.coverage(pk_h, ps_h, pe_h, strict=True, out=out) == oi.coverage(pk, ps, pe, strict=True)).all()
    left, valid, dist = gi.nearest_k1(pk_h, ps_h, pe_h)
    ol, ov, od = oi.nearest_k1(pk, ps, pe)
    assert (valid == ov).all() and (left[valid] == ol[ov]).all() and (dist[valid] == od[ov]).all()
    assert (gi.probe_arrow("count", *brg._cabi.key_runs(pk), host(ps.astype(np.int64)), host(pe.astype(np.uint64))) == oi.count_overlaps(pk, ps, pe)).all()
    gi.close()
    c2.close()


def test_gpu_nearest_k1_through_the_k_walker(ctx):
    """brg_nearest(k = 1) must behave as nearest_one (nearest_index.rs:110-115): a strict probe 1-2 bp wide is degenerate
    (qs + 1 > qe - 1), skips the overlap step (:222-225) and falls to the neighbour search"""
    rng = np.random.default_rng(21)
    nk = 3
    bk, bs, be = _rand(rng, 4000, nk, 0, 100_000, 300)
    pk = rng.integers(0, nk, 30_000).astype(np.uint32)
    ps = rng.integers(0, 100_000, 30_000).astype(np.int32)
    pe = (ps + rng.integers(0, 3, 30_000)).astype(np.int32)  # widths 0, 1, 2
    gi = ctx.build(bk, bs, be, nk)
    oi = O.OracleIndex(bk, bs, be, nk)
    for strict in (True, False):
        for inc in (True, False):
            g = gi.nearest_k(pk, ps, pe, 1, include_overlaps=inc, strict=strict)
            o = oi.nearest_k(pk, ps, pe, 1, include_overlaps=inc, strict=strict)
            assert len(g[0]) == len(o[0])
            for a, b in zip(g, o):
                assert (a == b).all()
            l1, v1, d1 = gi.nearest_k1(pk, ps, pe, include_overlaps=inc, strict=strict)
            assert (v1 == g[1]).all() and (l1[v1] == g[0][v1]).all() and (d1[v1] == g[3][v1]).all()
    gi.close()


def test_gpu_edge_cases(ctx):
    rng = np.random.default_rng(77)
    # empty build side, keys without rows, empty probe side
    z32, zu = np.zeros(0, np.int32), np.zeros(0, np.uint32)
    pk, ps, pe = _rand(rng, 1000, 4, 0, 5000, 50, absent_frac=0.2)
    _compare_all(ctx, zu, z32, z32, pk, ps, pe, 4)
    bk, bs, be = _rand(rng, 3000, 2, 0, 5000, 80)
    bk[:] = 3  # only key 3 of 6 has rows
    _compare_all(ctx, bk, bs, be, pk, ps, pe, 6)
    gi = ctx.build(bk, bs, be, 6)
    assert gi.count_overlaps(zu, z32, z32).shape == (0,)
    assert gi.overlap_pairs(zu, z32, z32)[0].shape == (0,)
    # duplicates, point intervals, many identical rows
    bs = np.repeat(np.array([10, 10, 10, 500, 500, 7], np.int32), 50)
    be = np.repeat(np.array([10, 20, 10, 900, 500, 7], np.int32), 50)
    bk = np.zeros(len(bs), np.uint32)
    pk = np.zeros(400, np.uint32)
    ps = rng.integers(0, 1000, 400).astype(np.int32)
    pe = (ps + rng.integers(0, 30, 400)).astype(np.int32)
    _compare_all(ctx, bk, bs, be, pk, ps, pe, 1, ks=(3, 60, 400))
    # degenerate probes after the strict adjustment (length 0/1 probes): count is pinned to 0 by the guard, nearest
    # falls to the neighbour search; coverage/pairs use the unguarded closed predicate (declared unpinned)
    bk, bs, be = _rand(rng, 2000, 1, 0, 3000, 100)
    pk = np.zeros(1500, np.uint32)
    ps = rng.integers(0, 3000, 1500).astype(np.int32)
    pe = (ps + rng.integers(-3, 2, 1500)).astype(np.int32)
    _compare_all(ctx, bk, bs, be, pk, ps, pe, 1)
    # extreme coordinates (wrapping i32 arithmetic must match the reference's release-mode behaviour)
    ext = np.array([-2**31, -2**31 + 1, -5, 0, 5, 2**31 - 2, 2**31 - 1], np.int64)
    bs = rng.choice(ext, 300)
    be = np.maximum(bs, rng.choice(ext, 300))
    ps = rng.choice(ext, 300)
    pe = np.maximum(ps, rng.choice(ext, 300))
    _compare_all(ctx, np.zeros(300, np.uint32), bs.astype(np.int32), be.astype(np.int32), np.zeros(300, np.uint32),
                 ps.astype(np.int32), pe.astype(np.int32), 1)


def test_gpu_nested_long_intervals(ctx):
    """one chromosome-long interval + many short ones: pair enumeration must stay exact (over-scan case)"""
    rng = np.random.default_rng(5)
    bk, bs, be = _rand(rng, 4000, 1, 0, 10**6, 300)
    bs[0], be[0] = 0, 10**6 + 500
    pk, ps, pe = _rand(rng, 3000, 1, 0, 10**6, 200)
    _compare_all(ctx, bk, bs, be, pk, ps, pe, 1, ks=(2,))


def test_gpu_config1_count_overlaps_1m_x_100k(ctx):
    """BASELINE configs[0]: count_overlaps 1M reads x 100k targets, 1 contig -- full size, bit-exact vs the oracle"""
    rng = np.random.default_rng(42)
    bs = rng.integers(0, 2**28, 100_000).astype(np.int32)
    be = (bs + rng.integers(100, 10_000, 100_000)).astype(np.int32)
    ps = rng.integers(0, 2**28, 1_000_000).astype(np.int32)
    pe = (ps + 150).astype(np.int32)
    gi = ctx.build(None, bs, be, 1)
    oi = O.OracleIndex(None, bs, be, 1)
    for order in ("unsorted", "sorted"):
        if order == "sorted":
            o = np.argsort(ps, kind="stable")
            ps, pe = ps[o], pe[o]
        got = gi.count_overlaps(None, ps, pe)
        want = oi.count_overlaps(None, ps, pe, threads=O.max_threads())
        assert (got == want).all(), order
        assert (gi.coverage(None, ps, pe) == oi.coverage(None, ps, pe, threads=O.max_threads())).all()
        gl, gv, gd = gi.nearest_k1(None, ps, pe)
        ol, ov, od = oi.nearest_k1(None, ps, pe, threads=O.max_threads())
        assert (gl == ol).all() and (gv == ov).all() and (gd == od).all()


def test_gpu_pairs_chunking_max_rows(ctx):
    """low-memory style chunking (IJ:1466-1573): concatenating capped calls reproduces the uncapped result"""
    rng = np.random.default_rng(9)
    bk, bs, be = _rand(rng, 20000, 3, 0, 200000, 400)
    pk, ps, pe = _rand(rng, 5000, 3, 0, 200000, 300)
    gi = ctx.build(bk, bs, be, 3)
    fb, fp = gi.overlap_pairs(pk, ps, pe)
    done, bb, pp = 0, [], []
    while done < len(ps):
        b, p, used = gi.overlap_pairs(pk[done:], ps[done:], pe[done:], max_rows=1000, return_consumed=True)
        assert used >= 1 and (len(b) <= 1000 or used == 1)
        bb.append(b)
        pp.append(p + done)
        done += used
    assert (np.concatenate(bb) == fb).all() and (np.concatenate(pp) == fp).all()


def test_gpu_large_scale_properties(ctx):
    """size-independent properties at a size the oracle would not finish quickly (16M probes x 2M build, 24 contigs)"""
    import torch

    g = torch.Generator(device="cuda").manual_seed(43)
    nb, npb, nk = 2_000_000, 16_000_000, 24
    bk = torch.randint(0, nk, (nb,), device="cuda", generator=g, dtype=torch.int32)
    bs = torch.randint(0, 2**27, (nb,), device="cuda", generator=g, dtype=torch.int32)
    be = bs + torch.randint(1, 2000, (nb,), device="cuda", generator=g, dtype=torch.int32)
    pk = torch.randint(0, nk, (npb,), device="cuda", generator=g, dtype=torch.int32)
    ps = torch.randint(0, 2**27, (npb,), device="cuda", generator=g, dtype=torch.int32)
    pe = ps + 150
    ix = ctx.build_dev(bk, bs, be, nk)
    out = torch.empty(npb, dtype=torch.int64, device="cuda")
    ix.count_overlaps_dev(pk, ps, pe, out)
    torch.cuda.synchronize()
    # (1) checksum of checksums: sum of counts == number of pairs found by the independent pair pipeline
    offsets = torch.empty(npb + 1, dtype=torch.int64, device="cuda")
    ix.pairs_count_dev(pk, ps, pe, offsets)
    torch.cuda.synchronize()
    total_pairs = int(offsets[-1].item())
    assert int(out.sum().item()) == total_pairs
    # (2) order independence: a permutation of the probes permutes the counts
    perm = torch.randperm(npb, device="cuda", generator=g)
    out2 = torch.empty_like(out)
    ix.count_overlaps_dev(pk[perm].contiguous(), ps[perm].contiguous(), pe[perm].contiguous(), out2)
    torch.cuda.synchronize()
    assert torch.equal(out2, out[perm])
    # (3) a sample of rows against the oracle
    oi = O.OracleIndex(bk.cpu().numpy().astype(np.uint32), bs.cpu().numpy(), be.cpu().numpy(), nk)
    sel = torch.arange(0, npb, 64, device="cuda")
    want = oi.count_overlaps(pk[sel].cpu().numpy().astype(np.uint32), ps[sel].cpu().numpy(), pe[sel].cpu().numpy(),
                             threads=O.max_threads())
    assert (out[sel].cpu().numpy() == want).all()
    # (4) emitted pairs satisfy the predicate and are exactly `total_pairs` many
    bi = torch.empty(total_pairs, dtype=torch.int32, device="cuda")
    pi = torch.empty(total_pairs, dtype=torch.int32, device="cuda")
    ix.pairs_emit_dev(pk, ps, pe, offsets, bi, pi)
    torch.cuda.synchronize()
    b64, p64 = bi.long(), pi.long()
    assert bool(((bk[b64] == pk[p64]) & (bs[b64] <= pe[p64]) & (be[b64] >= ps[p64])).all())
    assert bool((p64[1:] >= p64[:-1]).all())  # probe-major
    # (5) the chunked pipeline (256-row chunk offsets, staged emission) produces the same pairs in the same order,
    #     on unsorted probes (all chunks deferred) and on coordinate-sorted probes (tiled kernels)
    for sort_first in (False, True):
        if sort_first:
            o = torch.argsort(pk.long() * (1 << 32) + ps.long())
            qk, qs, qe = pk[o].contiguous(), ps[o].contiguous(), pe[o].contiguous()
            ix.pairs_count_dev(qk, qs, qe, offsets)
            ix.pairs_emit_dev(qk, qs, qe, offsets, bi, pi)
            torch.cuda.synchronize()
            assert int(offsets[-1].item()) == total_pairs
        else:
            qk, qs, qe = pk, ps, pe
        cb = torch.empty((npb + 255) // 256 + 1, dtype=torch.int64, device="cuda")
        rr = torch.empty(npb, dtype=torch.int64, device="cuda")
        ix.pairs_chunk_offsets_dev(qk, qs, qe, cb, rr)
        torch.cuda.synchronize()
        assert int(cb[-1].item()) == total_pairs
        assert torch.equal(cb[:-1], offsets[:-1:256])
        bi2 = torch.full((total_pairs,), -1, dtype=torch.int32, device="cuda")
        pi2 = torch.full((total_pairs,), -1, dtype=torch.int32, device="cuda")
        ix.pairs_fill_dev(qs, cb, rr, bi2, pi2)
        torch.cuda.synchronize()
        assert torch.equal(bi2, bi) and torch.equal(pi2, pi)
    # (6) device entry points accept columns that are not 16-byte aligned (views at odd offsets): scalar paths
    m = 1_000_001
    qk, qs, qe = pk[1:m + 1], ps[1:m + 1], pe[1:m + 1]
    o1 = torch.empty(m + 1, dtype=torch.int64, device="cuda")[1:]
    ix.count_overlaps_dev(qk, qs, qe, o1)
    torch.cuda.synchronize()
    assert torch.equal(o1, out[1:m + 1])
    cb = torch.empty((m + 255) // 256 + 1, dtype=torch.int64, device="cuda")
    rr = torch.empty(m + 1, dtype=torch.int64, device="cuda")[1:]
    ix.pairs_chunk_offsets_dev(qk, qs, qe, cb, rr)
    torch.cuda.synchronize()
    tot = int(cb[-1].item())
    assert tot == int(out[1:m + 1].sum().item())
    b3 = torch.empty(tot, dtype=torch.int32, device="cuda")
    p3 = torch.empty(tot, dtype=torch.int32, device="cuda")
    ix.pairs_fill_dev(qs, cb, rr, b3, p3)
    torch.cuda.synchronize()
    b64, p64 = b3.long(), p3.long()
    assert bool(((bk[b64] == qk[p64]) & (bs[b64] <= qe[p64]) & (be[b64] >= qs[p64])).all())


def test_gpu_index_export_import_roundtrip(ctx):
    """replication path used for multi-GPU: export header + slab, import into a fresh index, same answers"""
    import torch

    rng = np.random.default_rng(3)
    bk, bs, be = _rand(rng, 50000, 5, 0, 10**6, 500)
    pk, ps, pe = _rand(rng, 20000, 5, 0, 10**6, 200)
    ix = ctx.build(bk, bs, be, 5)
    header, slab = ix.export()
    ix2, slab2 = ctx.import_index(header)
    slab2.copy_(slab)
    torch.cuda.synchronize()
    assert (ix.count_overlaps(pk, ps, pe) == ix2.count_overlaps(pk, ps, pe)).all()
    assert (ix.coverage(pk, ps, pe) == ix2.coverage(pk, ps, pe)).all()
    a, b = ix.nearest_k1(pk, ps, pe), ix2.nearest_k1(pk, ps, pe)
    assert all((x == y).all() for x, y in zip(a, b))


def test_gpu_concurrent_probes_share_one_index(ctx):
    """the reference shares one build result across partitions polled on several threads (joins/utils.rs:73-107,
    count_overlaps.rs:127-153): probe entry points are re-entrant against an immutable index"""
    import threading

    rng = np.random.default_rng(12)
    nk = 4
    bk, bs, be = _rand(rng, 80_000, nk, 0, 5_000_000, 900)
    gi = ctx.build(bk, bs, be, nk)
    oi = O.OracleIndex(bk, bs, be, nk)
    parts = [_sorted_probes(np.random.default_rng(100 + t), 300_000 + 1000 * t, nk, 5_000_000, 1, 700) for t in range(6)]
    want = [(oi.count_overlaps(*p, threads=O.max_threads()), oi.coverage(*p, threads=O.max_threads()), oi.nearest_k1(*p, threads=O.max_threads()))
            for p in parts]
    got, errs = [None] * len(parts), []

    def work(t):
        try:
            for _ in range(3):
                got[t] = (gi.count_overlaps(*parts[t]), gi.coverage(*parts[t]), gi.nearest_k1(*parts[t]),
                          gi.overlap_pairs(*parts[t]))
        except Exception as ex:  # pragma: no cover
            errs.append(ex)

    th = [threading.Thread(target=work, args=(t,)) for t in range(len(parts))]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for t, p in enumerate(parts):
        assert (got[t][0] == want[t][0]).all() and (got[t][1] == want[t][1]).all()
        assert all((a == b).all() for a, b in zip(got[t][2], want[t][2]))
        gb, gp = O.canonical_pairs(*got[t][3])
        ob, op = O.canonical_pairs(*oi.overlap_pairs(*p))
        assert (gb == ob).all() and (gp == op).all()
    gi.close()


def test_gpu_take_dev_matches_index_select(ctx):
    """device-side `take` of fixed-width columns by the emitted pair indices (interval_join.rs:1655-1667)"""
    import torch

    g = torch.Generator(device="cuda").manual_seed(5)
    n_src, n = 1_000_003, 3_000_001
    idx = torch.randint(0, n_src, (n,), device="cuda", generator=g, dtype=torch.int32)
    valid = torch.rand(n, device="cuda", generator=g) > 0.1
    pad = (-n) % 32
    vb = torch.cat([valid, torch.zeros(pad, dtype=torch.bool, device="cuda")]).view(-1, 32).to(torch.int64)
    bits = (vb << torch.arange(32, device="cuda")).sum(1).to(torch.int32)  # LSB-first validity words
    for dtype in (torch.uint8, torch.int16, torch.int32, torch.int64):
        src = torch.randint(0, 100, (n_src,), device="cuda", generator=g).to(dtype)
        out = torch.empty(n, dtype=dtype, device="cuda")
        ctx.take_dev(src, idx, out)
        torch.cuda.synchronize()
        assert torch.equal(out, src[idx.long()])
        ctx.take_dev(src, idx, out, valid_bits=bits)
        torch.cuda.synchronize()
        assert torch.equal(out, torch.where(valid, src[idx.long()], torch.zeros((), dtype=dtype, device="cuda")))
