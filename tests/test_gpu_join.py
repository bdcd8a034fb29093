"""The one-launch pair join (brg_pairs_join_dev, csrc/brg_join.cu) against the per-row count -> scan -> emit pipeline and against the
oracle's order-independent checksum of the pair set (interval_join.rs:849-863 semantics: b.start <= q.end && b.end >= q.start,
unguarded).  BRG_JOIN_ORDERED: the same pairs in the same order (probe-major, matches of a probe in (start, end, row id) order).
Default (any order): each chunk's pairs contiguous at the slot its chunk-slot record names, in that same order inside the chunk."""
import zlib

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch

    import bio_ranges_b200 as brg

    c = brg.Context(0)
    yield torch, brg, c
    c.close()


def _dev(torch, a, dtype=None, offset=0):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32 else np.ascontiguousarray(a)).cuda()
    if offset:  # a view that starts `offset` elements into an allocation: not 16-byte aligned
        pad = torch.empty(t.numel() + offset, dtype=t.dtype, device="cuda")
        pad[offset:] = t
        t = pad[offset:]
    return t


def _reference_pairs(torch, ix, pk, ps, pe):
    n = ps.numel()
    offsets = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    ix.pairs_count_dev(pk, ps, pe, offsets)
    total = int(offsets[-1].item())
    bi = torch.empty(max(total, 1), dtype=torch.int32, device="cuda")
    pi = torch.empty(max(total, 1), dtype=torch.int32, device="cuda")
    if total:
        ix.pairs_emit_dev(pk, ps, pe, offsets, bi, pi)
    return total, bi[:total], pi[:total], offsets


def _join(torch, ix, pk, ps, pe, capacity, offset=0, ordered=True, slots=False):
    import bio_ranges_b200 as brg

    alloc = max(capacity, 1) + offset + 64
    bi = torch.full((alloc,), -7, dtype=torch.int32, device="cuda")
    pi = torch.full((alloc,), -7, dtype=torch.int32, device="cuda")
    total = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    rows = brg._cabi.lib().brg_pairs_join_chunk_rows()
    n_chunks = (ps.numel() + rows - 1) // rows
    cs = torch.full((2 * n_chunks,), -1, dtype=torch.int64, device="cuda") if slots else None
    ix.pairs_join_dev(pk, ps, pe, bi[offset:offset + capacity], pi[offset:offset + capacity], total, ordered=ordered, chunk_slots=cs)
    torch.cuda.synchronize()
    # nothing outside [offset, offset + capacity) was written
    assert (bi[:offset] == -7).all() and (bi[offset + capacity:] == -7).all()
    assert (pi[:offset] == -7).all() and (pi[offset + capacity:] == -7).all()
    if slots:
        return int(total.item()), bi[offset:offset + capacity], pi[offset:offset + capacity], cs.view(-1, 2), rows
    return int(total.item()), bi[offset:offset + capacity], pi[offset:offset + capacity]


def _check_any_order(torch, roff, rb, rp, total, got, gb, gp, cs, rows):
    """any-order output: exact total; the chunk-slot table tiles [0, total) without overlap; each chunk's pairs are the
    reference's pairs of its probe rows, in the reference's order"""
    assert got == total
    n = roff.numel() - 1
    bounds = torch.arange(0, n + rows, rows, device="cuda").clamp_(max=n)
    want_len = roff[bounds[1:]] - roff[bounds[:-1]]
    assert torch.equal(cs[:, 1], want_len)
    order = torch.argsort(cs[:, 0], stable=True)
    nz = order[cs[order, 1] > 0]
    starts, lens = cs[nz, 0], cs[nz, 1]
    assert total == 0 or (int(starts[0]) == 0 and torch.equal(starts[1:], (starts + lens)[:-1]) and int(starts[-1] + lens[-1]) == total)
    # gather the reference pairs into the layout the join chose and compare all of them at once
    src0 = roff[bounds[:-1]]
    dst = torch.repeat_interleave(cs[:, 0] - src0, want_len) + torch.arange(total, device="cuda")
    assert torch.equal(gb[dst], rb) and torch.equal(gp[dst], rp)


def _check(env, bk, bs, be, nk, pk, ps, pe, oracle=True, offsets=(0,)):
    torch, brg, ctx = env
    ix = ctx.build(bk, bs, be, nk, want=brg.WANT_PAIRS)
    for off in offsets:
        dk, ds, de = _dev(torch, pk, offset=off), _dev(torch, ps, offset=off), _dev(torch, pe, offset=off)
        total, rb, rp, roff = _reference_pairs(torch, ix, dk, ds, de)
        got, gb, gp = _join(torch, ix, dk, ds, de, total, offset=off)
        assert got == total
        assert torch.equal(gb, rb) and torch.equal(gp, rp)
        got, gb, gp, cs, rows = _join(torch, ix, dk, ds, de, total, offset=off, ordered=False, slots=True)
        _check_any_order(torch, roff, rb, rp, total, got, gb, gp, cs, rows)
        if total > 3:  # a buffer that is too small: exact total, the first `cap` pairs, nothing written past it
            cap = total // 3
            got2, gb2, gp2 = _join(torch, ix, dk, ds, de, cap, offset=off)
            assert got2 == total and torch.equal(gb2, rb[:cap]) and torch.equal(gp2, rp[:cap])
        got0, _, _ = _join(torch, ix, dk, ds, de, 0)
        assert got0 == total
        got0, _, _ = _join(torch, ix, dk, ds, de, 0, ordered=False)
        assert got0 == total
    if oracle:
        oi = O.OracleIndex(bk, bs, be, nk)
        want = oi.overlap_checksum(pk, ps, pe, threads=O.max_threads())
        assert O.pair_checksum(rb.cpu().numpy().view(np.uint32), rp.cpu().numpy().view(np.uint32)) == want
        # and through the host-buffer entry point in its any-order mode (the one-launch join)
        hb, hp = ix.overlap_pairs(pk, ps, pe, mode=brg.PAIRS_INNER_ANY_ORDER)
        assert len(hb) == total and O.pair_checksum(hb, hp) == want
        assert (np.bincount(hp, minlength=len(ps)) == np.diff(roff.cpu().numpy())).all()
    ix.close()


def _table(rng, n, nk, span, len_lo, len_hi, order="sorted"):
    k = rng.integers(0, nk, n).astype(np.uint32)
    s = rng.integers(0, span, n).astype(np.int64)
    e = s + rng.integers(len_lo, len_hi + 1, n)
    if order == "sorted":
        o = np.lexsort((s, k))
        k, s, e = k[o], s[o], e[o]
    return k, s.astype(np.int32), e.astype(np.int32)


REGIMES = {
    # name: (n_build, n_probe, n_keys, span, build len lo/hi, probe len lo/hi, order)
    "dense_sorted_variable": (100_000, 100_003, 3, 10_000_000, 1, 1550, 1, 1550, "sorted"),
    "dense_sorted_fixed": (100_000, 65_536, 2, 8_000_000, 1, 900, 150, 150, "sorted"),
    "sparse_build": (10_000, 120_000, 4, 12_000_000, 50, 2000, 1, 400, "sorted"),
    "build_much_denser": (400_000, 8_191, 1, 400_000, 1, 30, 1, 100, "sorted"),   # windows above the staged capacity
    "long_probes": (60_000, 30_000, 2, 2_000_000, 1, 200, 1, 60_000, "sorted"),     # candidate counts above the slot map
    "long_build_rows": (50_000, 50_000, 2, 5_000_000, 1, 100_000, 1, 100, "sorted"),  # prefix_max_end reaches far back
    "many_keys": (80_000, 80_000, 700, 200_000, 1, 300, 1, 300, "sorted"),          # most chunks span several keys
    "shuffled": (60_000, 40_000, 3, 3_000_000, 1, 900, 1, 900, "shuffled"),
    "tiny": (7, 5, 1, 100, 1, 30, 1, 30, "sorted"),
    "one_chunk_minus_one": (3_000, 255, 1, 100_000, 1, 300, 1, 300, "sorted"),
}


@pytest.mark.parametrize("name", list(REGIMES))
def test_join_regimes(env, name):
    nb, npb, nk, span, bl0, bl1, pl0, pl1, order = REGIMES[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    bk, bs, be = _table(rng, nb, nk, span, bl0, bl1, "shuffled")
    pk, ps, pe = _table(rng, npb, nk, span, pl0, pl1, order)
    _check(env, bk, bs, be, nk, pk, ps, pe, offsets=(0, 1) if name in ("dense_sorted_variable", "tiny") else (0,))


def test_join_edge_rows(env):
    """absent keys (whole chunks and single rows), degenerate probes (end < start), build rows with start > end, i32 extremes"""
    rng = np.random.default_rng(11)
    nk = 3
    bk, bs, be = _table(rng, 30_000, nk, 1_000_000, 1, 500, "shuffled")
    inv = rng.random(len(bs)) < 0.01
    bs2 = np.where(inv, be, bs).astype(np.int32)
    be2 = np.where(inv, bs, be).astype(np.int32)
    bs2[:3] = [-2**31, 2**31 - 1, -2**31]
    be2[:3] = [-2**31, 2**31 - 1, 2**31 - 1]
    pk, ps, pe = _table(rng, 40_000, nk, 1_000_000, 0, 700, "sorted")
    pk[1000:1900] = O.ABSENT          # a run of absent keys covering whole chunks
    pk[rng.random(len(pk)) < 0.01] = O.ABSENT
    deg = rng.random(len(ps)) < 0.02
    pe = np.where(deg, ps - rng.integers(1, 50, len(ps)), pe).astype(np.int32)
    ps[:2] = -2**31
    pe[:2] = [-2**31, 2**31 - 1]
    ps[-1], pe[-1] = 2**31 - 1, 2**31 - 1
    for b_s, b_e in ((bs, be), (bs2, be2)):
        _check(env, bk, b_s, b_e, nk, pk, ps, pe)


def test_join_piles(env):
    """piles of duplicate build rows under wide probes: one row alone exceeds the slot map (the warp sweeps it), chunks whose
    candidates exceed the staging capacity (count pass + direct emission)"""
    rng = np.random.default_rng(5)
    bs = np.concatenate([np.full(9000, 5000), rng.integers(0, 100_000, 20_000)]).astype(np.int32)
    be = (bs + np.concatenate([np.full(9000, 10), rng.integers(1, 300, 20_000)])).astype(np.int32)
    bk = np.zeros(len(bs), np.uint32)
    pk, ps, pe = _table(rng, 5_000, 1, 100_000, 1, 400, "sorted")
    _check(env, bk, bs, be, 1, pk, ps, pe)


def test_join_empty(env):
    torch, brg, ctx = env
    z = np.zeros(0, np.int32)
    ix = ctx.build(np.zeros(4, np.uint32), np.arange(4, dtype=np.int32), np.arange(4, dtype=np.int32) + 5, 1, want=brg.WANT_PAIRS)
    e = torch.empty(0, dtype=torch.int32, device="cuda")
    got, _, _ = _join(torch, ix, e, e, e, 16)
    assert got == 0
    hb, hp = ix.overlap_pairs(np.zeros(0, np.uint32), z, z, mode=brg.PAIRS_INNER_ANY_ORDER)
    assert len(hb) == 0
    ix.close()


def test_join_repeatable_under_load(env):
    """the look-back is order-dependent machinery: many launches back to back give the same pairs every time"""
    torch, brg, ctx = env
    rng = np.random.default_rng(3)
    bk, bs, be = _table(rng, 300_000, 5, 30_000_000, 1, 1550, "shuffled")
    pk, ps, pe = _table(rng, 1_000_003, 5, 30_000_000, 1, 1550, "sorted")
    ix = ctx.build(bk, bs, be, 5, want=brg.WANT_PAIRS)
    dk, ds, de = _dev(torch, pk), _dev(torch, ps), _dev(torch, pe)
    total, rb, rp, _ = _reference_pairs(torch, ix, dk, ds, de)
    bi = torch.empty(total, dtype=torch.int32, device="cuda")
    pi = torch.empty(total, dtype=torch.int32, device="cuda")
    tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    for _ in range(20):
        bi.fill_(-1)
        ix.pairs_join_dev(dk, ds, de, bi, pi, tot, ordered=True)
        assert int(tot.item()) == total and torch.equal(bi, rb) and torch.equal(pi, rp)
    want = O.pair_checksum(rb.cpu().numpy().view(np.uint32), rp.cpu().numpy().view(np.uint32))
    for _ in range(5):
        bi.fill_(-1)
        ix.pairs_join_dev(dk, ds, de, bi, pi, tot)
        assert int(tot.item()) == total
        assert O.pair_checksum(bi.cpu().numpy().view(np.uint32), pi.cpu().numpy().view(np.uint32)) == want
    ix.close()
