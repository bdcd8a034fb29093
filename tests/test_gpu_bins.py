"""The bins tables behind the one-row-per-thread kernels (unsorted probe tables), on the GPU against the oracle:
count bins over (start, end), coverage bins over the merged intervals, nearest / pair bins over (start, prefix-max end),
each in the wide (12 x u16) and the packed (18 x 10 bit) record format (csrc/brg_bins.cuh), on build sides that make the
records overflow (piles, deep nesting), that cannot have the nearest / pair bins at all (rows with start > end) and that are too
sparse for the packed format.  BRG_CB_PACK / BRG_NB_BINS are read at every build, so one process covers all combinations."""
import numpy as np
import pytest

from oracle import oracle as O
from test_gpu_parity import _compare_all, _rand


def _compare_some(ctx, bk, bs, be, pk, ps, pe, nk, others=True, pair_rows=None):
    """count_overlaps on every probe row; coverage / nearest k = 1 unless `others` is off; pairs on the first `pair_rows`
    rows (all when None)"""
    gi = ctx.build(bk, bs, be, nk)
    oi = O.OracleIndex(bk, bs, be, nk)
    for strict in (False, True):
        assert (gi.count_overlaps(pk, ps, pe, strict=strict) == oi.count_overlaps(pk, ps, pe, strict=strict)).all()
        if not others:
            continue
        assert (gi.coverage(pk, ps, pe, strict=strict) == oi.coverage(pk, ps, pe, strict=strict)).all()
        for inc in (True, False):
            gl, gv, gd = gi.nearest_k1(pk, ps, pe, include_overlaps=inc, strict=strict)
            ol, ov, od = oi.nearest_k1(pk, ps, pe, include_overlaps=inc, strict=strict)
            assert (gv == ov).all() and (gl == ol).all() and (gd == od).all()
    m = len(ps) if pair_rows is None else pair_rows
    gb, gp = O.canonical_pairs(*gi.overlap_pairs(pk[:m], ps[:m], pe[:m]))
    ob, op = O.canonical_pairs(*oi.overlap_pairs(pk[:m], ps[:m], pe[:m]))
    assert len(gb) == len(ob) and (gb == ob).all() and (gp == op).all()
    gi.close()

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import bio_ranges_b200 as brg

    c = brg.Context(0)
    yield c
    c.close()


def _dense(rng):
    return _rand(rng, 40_000, 3, 0, 700_000, 40) + (3, True, True)


def _dense_long(rng):  # long build rows: many intervals open at every bin (depth), prefix-max ends far ahead of the starts
    return _rand(rng, 30_000, 2, 0, 400_000, 8_000) + (2, True, True)


def _piles(rng):  # a third of the rows piled on a few coordinates: those records overflow and the rows fall back to the directories
    k, s, e = _rand(rng, 30_000, 2, 0, 400_000, 200)
    hot = rng.integers(0, 400_000, 12)
    m = rng.random(len(s)) < 0.33
    s[m] = hot[rng.integers(0, len(hot), int(m.sum()))]
    e[m] = s[m] + rng.integers(0, 50, int(m.sum()))
    return k, s, e, 2, True, True


def _deep(rng):  # more than 65535 intervals open at once: the depth field overflows (a probe in there meets 10^5 build rows:
    n = 120_000    # the test joins only a few dozen probe rows)
    s = rng.integers(0, 100_000, n).astype(np.int32)
    e = (s.astype(np.int64) + 500_000 + rng.integers(0, 1000, n)).astype(np.int32)
    return np.zeros(n, np.uint32), s, e, 1, True, True


def _inverted(rng):  # rows with start > end: no nearest / pair bins; count records with more ends than starts before them overflow
    # (count_overlaps and the pair join are defined for such rows -- interval_tree.rs:46-48, interval_join.rs:849-863 -- and are
    #  what this side checks; coverage and nearest over them are not pinned by any reference test and are left out)
    k, s, e = _rand(rng, 40_000, 3, 0, 500_000, 300)
    m = rng.random(len(s)) < 0.05
    s[m], e[m] = e[m] + 1, s[m].copy()
    return k, s, e, 3, True, False


def _sparse(rng):  # bins wider than 1024: the packed format does not apply
    return _rand(rng, 5_000, 2, 0, 2_000_000_000, 5_000) + (2, False, True)


BUILD_SIDES = {"dense": _dense, "dense_long": _dense_long, "piles": _piles, "deep": _deep, "inverted": _inverted, "sparse": _sparse}


@pytest.mark.parametrize("side", list(BUILD_SIDES), ids=list(BUILD_SIDES))
@pytest.mark.parametrize("pack,nbins", [("1", "1"), ("0", "1"), ("0", "0")], ids=["packed-nearest_bins", "wide-nearest_bins", "wide-directories"])
def test_gpu_bins_tables(ctx, monkeypatch, pack, nbins, side):
    monkeypatch.setenv("BRG_CB_PACK", pack)
    monkeypatch.setenv("BRG_NB_BINS", nbins)
    rng = np.random.default_rng({"dense": 1, "dense_long": 2, "piles": 3, "deep": 4, "inverted": 5, "sparse": 6}[side])
    bk, bs, be, nk, packable, has_nb = BUILD_SIDES[side](rng)
    ix = ctx.build(bk, bs, be, nk)
    info = [ix.bins_info(t) for t in range(3)]
    ix.close()
    assert info[0]["records"] > 0
    want_fmt = 1 if (pack == "1" and packable) else 0
    assert info[0]["format"] == want_fmt, info
    assert info[0]["bin_shift"] <= (10 if want_fmt else 16)
    assert (info[2]["records"] > 0) == (nbins == "1" and has_nb), info
    if info[2]["records"]:
        assert info[2]["format"] == want_fmt, info
    # unsorted probe rows (more than one 256-row chunk: tiled kernel -> deferred chunks -> per-row kernels), a few absent keys,
    # probes inside, around and far outside the build side's coordinate range, 1-2 bp probes (degenerate under strict)
    lo, hi = int(bs.min()) - 5_000, int(max(bs.max(), be.max())) + 5_000
    pk, ps, pe = _rand(rng, 3_000, nk, lo, hi, 600, absent_frac=0.03)
    pk2, ps2, pe2 = _rand(rng, 800, nk, lo, hi, 3)
    pk3, ps3, pe3 = _rand(rng, 300, nk, lo - 3_000_000, hi + 3_000_000, 100_000)
    pk, ps, pe = np.concatenate([pk, pk2, pk3]), np.concatenate([ps, ps2, ps3]), np.concatenate([pe, pe2, pe3])
    if side == "deep":
        _compare_some(ctx, bk, bs, be, pk, ps, pe, nk, pair_rows=40)
    elif side == "inverted":
        _compare_some(ctx, bk, bs, be, pk, ps, pe, nk, others=False)
    else:
        _compare_all(ctx, bk, bs, be, pk, ps, pe, nk, ks=(2,))


def test_gpu_pair_bins_on_request(ctx, monkeypatch):
    """a pairs-only index has the nearest / pair bins only under BRG_NB_BINS=2; the join is the same either way"""
    import bio_ranges_b200 as brg

    rng = np.random.default_rng(11)
    bk, bs, be = _rand(rng, 30_000, 2, 0, 500_000, 300)
    pk, ps, pe = _rand(rng, 5_000, 2, -2_000, 505_000, 400, absent_frac=0.02)
    got = []
    for v, has in (("1", False), ("2", True)):
        monkeypatch.setenv("BRG_NB_BINS", v)
        ix = ctx.build(bk, bs, be, 2, brg.WANT_PAIRS)
        assert (ix.bins_info(2)["records"] > 0) == has
        got.append(O.canonical_pairs(*ix.overlap_pairs(pk, ps, pe)))
        ix.close()
    want = O.canonical_pairs(*O.OracleIndex(bk, bs, be, 2).overlap_pairs(pk, ps, pe))
    for g in got:
        assert len(g[0]) == len(want[0]) and (g[0] == want[0]).all() and (g[1] == want[1]).all()
