"""Seeded fuzz of the CUDA path against the oracle on small tables built from a SMALL coordinate alphabet: many ties,
duplicates, zero-width and touching intervals, i32 extremes, sorted and unsorted probe tables, full and partial 256-row
chunks.  Every function of the hot path, strict and weak (count / coverage / nearest k=1 / nearest k / pairs / semi / anti)."""
import numpy as np
import pytest

from oracle import oracle as O
from test_gpu_parity import _compare_all

pytestmark = pytest.mark.gpu

I32_MIN, I32_MAX = -2**31, 2**31 - 1


@pytest.fixture(scope="module")
def ctx():
    import bio_ranges_b200 as brg

    return brg.default_context(0)


def _case(seed):
    rng = np.random.default_rng(10_000 + seed)
    nk = int(rng.choice([1, 2, 5]))
    nb = int(rng.choice([1, 3, 40, 700, 5000]))
    npb = int(rng.choice([1, 8, 255, 256, 257, 512, 777, 2048, 4099]))
    # coordinate alphabet: a few dozen values, optionally with the i32 extremes (kept one step inside for the strict +-1)
    alphabet = np.unique(rng.integers(-50, int(rng.choice([60, 3000, 2_000_000])), int(rng.choice([6, 40, 400]))))
    if seed % 5 == 0:
        alphabet = np.unique(np.concatenate([alphabet, [I32_MIN + 2, I32_MAX - 2]]))
    alphabet = alphabet.astype(np.int64)

    def table(n, zero_width_p):
        k = rng.integers(0, nk, n).astype(np.uint32)
        a, b = rng.choice(alphabet, n), rng.choice(alphabet, n)
        s, e = np.minimum(a, b), np.maximum(a, b)
        z = rng.random(n) < zero_width_p
        e = np.where(z, s, e)
        return k, s, e

    bk, bs, be = table(nb, 0.2)
    pk, ps, pe = table(npb, 0.2)
    if seed % 3 == 0:  # a few probe keys that the build side does not have
        pk[rng.random(npb) < 0.05] = O.ABSENT
    if seed % 2 == 0:  # coordinate-sorted probe table (the tiled kernels), else shuffled (the per-row kernels)
        o = np.lexsort((ps, pk))
        pk, ps, pe = pk[o], ps[o], pe[o]
    return nk, (bk, bs.astype(np.int32), be.astype(np.int32)), (pk, ps.astype(np.int32), pe.astype(np.int32))


@pytest.mark.parametrize("seed", range(150))
def test_gpu_fuzz_small_alphabet(ctx, seed):
    nk, (bk, bs, be), (pk, ps, pe) = _case(seed)
    _compare_all(ctx, bk, bs, be, pk, ps, pe, nk, ks=(2,))
