"""The identities the CUDA kernels lean on, checked on the CPU with numpy (no GPU, no library): they are statements about sorted
arrays, so a brute-force count settles them.

 * count identity (interval_tree.rs:41-49): overlaps = #(start <= qe) - #(end < qs) for build rows with start <= end;
 * overlap test of nearest / pairs (nearest_index.rs:222-234): a row of the first n_le = #(start <= qe) rows in start order
   overlaps iff the first index whose prefix-max end >= qs lies below n_le, and no row before that index overlaps;
 * the nearest bins' shortcut (csrc/brg_probe.cu, nearest_one_bins): without an overlap, for a proper probe and build rows with
   start <= end, #(end < qs) == #(start <= qe) -- the left neighbour is row n_le - 1 of the end-sorted order."""
import numpy as np
import pytest


def _build(rng, n, span, maxlen):
    s = np.sort(rng.integers(0, span, n))
    e = s + rng.integers(0, maxlen, n)
    o = np.lexsort((e, s))
    return s[o], e[o]


@pytest.mark.parametrize("seed", range(6))
def test_rank_identities(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([1, 7, 300, 4000]))
    span = int(rng.choice([50, 5_000, 1_000_000]))
    s, e = _build(rng, n, span, int(rng.choice([1, 30, 2_000])))
    pme = np.maximum.accumulate(e)
    e_sorted = np.sort(e)
    qs = rng.integers(-100, span + 2_100, 3_000)
    qe = qs + rng.integers(0, 400, 3_000)
    n_le = np.searchsorted(s, qe, side="right")          # #(start <= qe)
    n_lt = np.searchsorted(e_sorted, qs, side="left")    # #(end < qs)
    i_pme = np.searchsorted(pme, qs, side="left")        # first row whose prefix-max end >= qs
    brute = ((s[None, :] <= qe[:, None]) & (e[None, :] >= qs[:, None]))
    cnt = brute.sum(1)
    assert (n_le - n_lt == cnt).all()
    has = cnt > 0
    assert ((i_pme < n_le) == has).all()
    first = np.where(has, brute.argmax(1), -1)
    assert (first[has] >= i_pme[has]).all()              # rows before i_pme end before qs
    # the row at i_pme is where the prefix maximum first reaches qs, so its own end does, and it starts by qe (i_pme < n_le):
    assert (first[has] == i_pme[has]).all()              # first_overlap_by_start is row i_pme
    assert (n_lt[~has] == n_le[~has]).all()              # the shortcut of the nearest bins
