"""Pins the CPU oracle (oracle/brg_oracle.c) to the reference: unit-test literals, CSV expected tables, the
expected_coverage.parquet golden file, the reference-native superintervals.hpp (oracle/_ref) and brute force."""
import numpy as np
import pytest

import reference_cases as rc
from oracle import oracle as O


def make_index(key, start, end, n_keys):
    return O.OracleIndex(key, start, end, n_keys)


@pytest.mark.parametrize("case", rc.ALL_CASES, ids=lambda c: c.__name__)
def test_oracle_reference_case(case):
    case(make_index)


def test_merge_intervals_literals():
    # IT:291-353
    assert O.merge_intervals([], [])[0].size == 0
    f, l = O.merge_intervals([1, 10, 20], [5, 15, 25])
    assert list(zip(f, l)) == [(1, 5), (10, 15), (20, 25)]
    f, l = O.merge_intervals([1, 3, 10], [5, 8, 15])
    assert list(zip(f, l)) == [(1, 8), (10, 15)]
    f, l = O.merge_intervals([1, 5], [5, 10])
    assert list(zip(f, l)) == [(1, 10)]
    f, l = O.merge_intervals([150, 190, 300, 500], [250, 300, 501, 700])
    assert list(zip(f, l)) == [(150, 700)]
    f, l = O.merge_intervals([10, 1, 3], [15, 5, 8])
    assert list(zip(f, l)) == [(1, 8), (10, 15)]
    # [1,5],[6,10] do NOT merge (6 <= 5 is false)
    f, l = O.merge_intervals([1, 6], [5, 10])
    assert list(zip(f, l)) == [(1, 5), (6, 10)]


def _rand(rng, n, n_keys, span, maxlen):
    k = rng.integers(0, n_keys, n).astype(np.uint32)
    s = rng.integers(-span, span, n).astype(np.int32)
    e = (s + rng.integers(0, maxlen, n)).astype(np.int32)
    return k, s, e


def _brute_pairs(bk, bs, be, pk, ps, pe):
    m = (bk[None, :] == pk[:, None]) & (bs[None, :] <= pe[:, None]) & (be[None, :] >= ps[:, None])
    p, b = np.nonzero(m)
    return b.astype(np.uint32), p.astype(np.uint32)


@pytest.mark.parametrize("seed", range(6))
def test_oracle_vs_bruteforce(seed):
    rng = np.random.default_rng(seed)
    nk = int(rng.integers(1, 5))
    bk, bs, be = _rand(rng, int(rng.integers(0, 400)), nk, 2000, int(rng.choice([2, 50, 3000])))
    pk, ps, pe = _rand(rng, 300, nk + 1, 2200, int(rng.choice([1, 80, 1000])))
    pk[pk == nk] = O.ABSENT
    ix = make_index(bk, bs, be, nk)
    # pairs == brute force (closed predicate)
    b, p = O.canonical_pairs(*ix.overlap_pairs(pk, ps, pe))
    bb, bp = O.canonical_pairs(*_brute_pairs(bk, bs, be, pk, ps, pe))
    assert (b == bb).all() and (p == bp).all()
    cnt = np.bincount(bp, minlength=len(ps))
    assert (ix.overlap_counts(pk, ps, pe) == cnt).all()
    for strict in (False, True):
        d = 1 if strict else 0
        # count_overlaps == predicate count whenever the (adjusted) probe is non-degenerate (IT:41-49)
        _, p2 = _brute_pairs(bk, bs, be, pk, ps + d, pe - d)
        want = np.bincount(p2, minlength=len(ps))
        want[(pe - d) < (ps + d)] = 0
        assert (ix.count_overlaps(pk, ps, pe, strict=strict) == want).all()
        # coverage == literal formula over brute-force merged intervals
        got = ix.coverage(pk, ps, pe, strict=strict)
        for i in range(0, len(ps), 7):
            acc = 0
            if pk[i] != O.ABSENT:
                sel = bk == pk[i]
                mf, ml = O.merge_intervals(bs[sel], be[sel])
                qs, qe = int(ps[i]) + d, int(pe[i]) - d
                for f, l in zip(mf, ml):
                    if f <= qe and l >= qs:
                        acc += max(1, min(qe + 1, int(l)) - max(qs - 1, int(f)))
            assert got[i] == acc
        # nearest k=1 == exhaustive minimum under (distance, start, end, position) with the reference's asymmetry
        for inc in (True, False):
            left, valid, dist = ix.nearest_k1(pk, ps, pe, include_overlaps=inc, strict=strict, compute_distance=True)
            for i in range(0, len(ps), 5):
                sel = np.nonzero(bk == pk[i])[0] if pk[i] != O.ABSENT else np.zeros(0, int)
                qs, qe = int(ps[i]) + d, int(pe[i]) - d
                want = _nearest_exhaustive(bs, be, sel, qs, qe, inc)
                if want is None:
                    assert not valid[i]
                else:
                    assert valid[i] and left[i] == want, (i, qs, qe, inc)
                    rs, re_ = int(ps[i]), int(pe[i])
                    s_, e_ = int(bs[want]), int(be[want])
                    dd = s_ - re_ if re_ < s_ else (rs - e_ if e_ < rs else 0)
                    assert dist[i] == dd


def _nearest_exhaustive(bs, be, sel, qs, qe, include_overlaps):
    """Independent restatement of NI:91-101 / 192-234 without sorted arrays."""
    if len(sel) == 0:
        return None
    recs = [(int(bs[j]), int(be[j]), int(j)) for j in sel]
    if include_overlaps and qe >= qs:
        ov = [r for r in recs if r[0] <= qe and r[1] >= qs]
        if ov:
            return min(ov)[2]
    lefts = [r for r in recs if r[1] < qs]
    rights = [r for r in recs if r[0] > qe]
    left = max(lefts, key=lambda r: (r[1], r[0], r[2])) if lefts else None
    right = min(rights) if rights else None
    if left is None and right is None:
        return None
    if left is None:
        return right[2]
    if right is None:
        return left[2]

    def dist(r):
        return r[0] - qe if qe < r[0] else (qs - r[1] if r[1] < qs else 0)

    return left[2] if (dist(left), left) <= (dist(right), right) else right[2]


@pytest.mark.parametrize("seed", range(3))
def test_oracle_nearest_k_consistency(seed):
    """k>1: first element equals k=1 answer when overlaps are included; rows are unique per probe; <= k per probe."""
    rng = np.random.default_rng(100 + seed)
    bk, bs, be = _rand(rng, 200, 2, 1500, 60)
    pk, ps, pe = _rand(rng, 150, 2, 1600, 40)
    ix = make_index(bk, bs, be, 2)
    for inc in (True, False):
        l1, v1, _ = ix.nearest_k1(pk, ps, pe, include_overlaps=inc)
        left, valid, right, dist = ix.nearest_k(pk, ps, pe, k=4, include_overlaps=inc)
        assert valid.all()
        for i in range(len(ps)):
            rows = left[right == i]
            assert 1 <= len(rows) <= 4 and len(set(rows.tolist())) == len(rows)
            assert rows[0] == l1[i]
            d = dist[right == i]
            if not inc:
                assert (d > 0).all()


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libsi_ref.so not built (reference tree absent)")
@pytest.mark.parametrize("seed", range(3))
def test_oracle_vs_reference_superintervals(seed):
    """The reference's own vendored superintervals.hpp enumerates the same closed-interval pairs (TEST:93,189)."""
    rng = np.random.default_rng(200 + seed)
    bk, bs, be = _rand(rng, 5000, 3, 100000, int(rng.choice([50, 5000])))
    bs = np.abs(bs); be = (bs + (be - bs.astype(np.int64) * 0) * 0 + rng.integers(0, 500, len(bs))).astype(np.int32)
    pk, ps, pe = _rand(rng, 3000, 3, 100000, 300)
    ps = np.abs(ps); pe = (ps + rng.integers(0, 300, len(ps))).astype(np.int32)
    ix = make_index(bk, bs, be, 3)
    si = O.SuperIntervalsRef(bk, bs, be, 3)
    b, p = O.canonical_pairs(*ix.overlap_pairs(pk, ps, pe))
    rb, rp = O.canonical_pairs(*si.pairs(pk, ps, pe))
    assert len(b) == len(rb) and (b == rb).all() and (p == rp).all()
    assert (si.counts(pk, ps, pe) == ix.overlap_counts(pk, ps, pe)).all()


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libsi_ref.so not built (reference tree absent)")
def test_reference_superintervals_on_golden_fixture():
    from _fixtures import golden

    g = golden()
    nk = len(g["contigs"])
    si = O.SuperIntervalsRef(g["fbrain_contig"].astype(np.uint32), g["fbrain_start"], g["fbrain_end"], nk)
    ix = make_index(g["fbrain_contig"].astype(np.uint32), g["fbrain_start"], g["fbrain_end"], nk)
    pk = g["exons_contig"].astype(np.uint32)
    rb, rp = O.canonical_pairs(*si.pairs(pk, g["exons_start"], g["exons_end"]))
    b, p = O.canonical_pairs(*ix.overlap_pairs(pk, g["exons_start"], g["exons_end"]))
    assert len(rb) == 54343 and (rb == b).all() and (rp == p).all()


def test_oracle_partitioned_pair_join_materialises_the_same_pairs():
    """bench.py's CPU baseline of the pairs workload (brgo_overlap_pairs_partitions: every thread materialises the index vectors
    of its probe-row range, IJ:1616-1653) produces the pair set of the single-threaded join, whatever the partition count"""
    rng = np.random.default_rng(17)
    bk, bs, be = _rand(rng, 20_000, 3, 300_000, 400)
    pk, ps, pe = _rand(rng, 30_001, 3, 301_000, 300)
    oi = O.OracleIndex(bk, bs, be, 3)
    bi, pi = oi.overlap_pairs(pk, ps, pe)
    want = O.pair_checksum(bi, pi)
    assert want[3] == len(bi) > 0
    for threads in (1, 2, 7, 16):
        assert oi.overlap_pairs_partitions(pk, ps, pe, threads=threads, checksum=True) == want
        assert oi.overlap_pairs_partitions(pk, ps, pe, threads=threads) == len(bi)
    assert oi.overlap_checksum(pk, ps, pe, threads=4) == want
