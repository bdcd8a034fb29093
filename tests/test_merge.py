"""merge: the oracle pinned to the reference's merge tests (tests/integration_test.rs:2062-2410,
testing/data/merge/input.csv), the CUDA path against the oracle, and the table function through the host layer."""
import numpy as np
import pytest

from _fixtures import RANGES_READS

# testing/data/merge/input.csv (test data, 22 rows)
MERGE_INPUT = [
    ("chr1", 150, 250), ("chr1", 190, 300), ("chr1", 300, 501), ("chr1", 500, 700), ("chr1", 22000, 22300), ("chr1", 15000, 15001),
    ("chr2", 150, 250), ("chr2", 190, 300), ("chr2", 300, 500), ("chr2", 500, 700), ("chr2", 22000, 22300), ("chr2", 15000, 15001),
    ("chr1", 100, 190), ("chr1", 200, 290), ("chr1", 400, 600), ("chr1", 10000, 20000), ("chr1", 22100, 22101),
    ("chr2", 100, 190), ("chr2", 200, 290), ("chr2", 400, 600), ("chr2", 10000, 20000), ("chr2", 22100, 22101),
]
# TEST:2104-2116 (strict), TEST:2335-2346 (reads.csv, weak)
EXPECTED_BIOFRAME = [("chr1", 100, 300, 4), ("chr1", 300, 700, 3), ("chr1", 10000, 20000, 2), ("chr1", 22000, 22300, 2),
                     ("chr2", 100, 300, 4), ("chr2", 300, 700, 3), ("chr2", 10000, 20000, 2), ("chr2", 22000, 22300, 2)]
EXPECTED_READS = [("chr1", 150, 700, 4), ("chr1", 15000, 15000, 1), ("chr1", 22000, 22300, 1), ("chr2", 150, 700, 4),
                  ("chr2", 15000, 15000, 1), ("chr2", 22000, 22300, 1), ("chr3", 234, 300, 1)]
LITERALS = [  # (rows, min_dist, strict, expected)  TEST:2123-2404
    ([("a", 100, 200), ("a", 150, 250), ("a", 300, 400)], 0, False, [("a", 100, 250, 2), ("a", 300, 400, 1)]),
    ([("a", 100, 200), ("a", 201, 300)], 0, False, [("a", 100, 200, 1), ("a", 201, 300, 1)]),
    ([("a", 100, 200), ("a", 201, 300)], 1, False, [("a", 100, 300, 2)]),
    ([("a", 100, 300), ("a", 150, 250), ("a", 200, 400)], 0, False, [("a", 100, 400, 3)]),
    ([("a", 100, 200)], 0, False, [("a", 100, 200, 1)]),
    ([], 0, False, []),
    ([("a", 100, 150), ("a", 150, 200)], 0, True, [("a", 100, 150, 1), ("a", 150, 200, 1)]),
    ([("a", 100, 150), ("a", 150, 200)], 0, False, [("a", 100, 200, 2)]),
    (MERGE_INPUT, 0, True, EXPECTED_BIOFRAME),
    (RANGES_READS, 0, False, EXPECTED_READS),
]


def _encode_sorted(rows):
    names = sorted({r[0] for r in rows})
    ids = {n: i for i, n in enumerate(names)}
    return names, np.array([ids[r[0]] for r in rows], np.uint32), np.array([r[1] for r in rows], np.int64), np.array([r[2] for r in rows], np.int64)


def _decode(names, out):
    k, s, e, c = out
    return [(names[int(a)], int(b), int(cc), int(d)) for a, b, cc, d in zip(k, s, e, c)]


@pytest.mark.parametrize("case", range(len(LITERALS)))
def test_oracle_merge_reference_literals(case):
    from oracle import oracle as O

    rows, min_dist, strict, want = LITERALS[case]
    names, k, s, e = _encode_sorted(rows)
    got = _decode(names, O.merge(k, s, e, max(1, len(names)), min_dist, strict))
    assert got == want  # contigs in name order, runs in coordinate order


def _random_case(seed):
    rng = np.random.default_rng(seed)
    nk = int(rng.choice([1, 3, 30]))
    n = int(rng.choice([1, 40, 3000, 200_000]))
    span = int(rng.choice([500, 100_000, 2_000_000_000]))
    k = rng.integers(0, nk, n).astype(np.uint32)
    s = rng.integers(-span if seed % 2 else 0, span, n).astype(np.int64)
    e = np.minimum(s + rng.integers(0, int(rng.choice([1, 50, 5000])), n), 2**31 - 1)
    md = int(rng.choice([0, 0, 1, 17, 10_000, 2**40]))
    return nk, k, s, e, md, bool(seed % 3 == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_gpu_merge_vs_oracle(seed):
    import bio_ranges_b200 as brg
    from oracle import oracle as O

    nk, k, s, e, md, strict = _random_case(seed)
    ctx = brg.default_context(0)
    gk, gs, ge, gc = ctx.merge(k, s.astype(np.int32), e.astype(np.int32), nk, md, strict)
    ok, os_, oe, oc = O.merge(k, s, e, nk, md, strict)
    assert len(gk) == len(ok)
    assert (gk == ok).all() and (gs == os_).all() and (ge == oe).all() and (gc == oc).all()
    assert int(gc.sum()) == len(k)


@pytest.mark.gpu
def test_gpu_merge_rejects_inverted_intervals():
    import bio_ranges_b200 as brg

    ctx = brg.default_context(0)
    with pytest.raises(brg.BioRangesError) as ei:
        ctx.merge(np.zeros(3, np.uint32), np.array([10, 50, 70], np.int32), np.array([20, 40, 80], np.int32), 1)
    assert ei.value.status == 5 and "start > end" in str(ei.value)


@pytest.mark.gpu
def test_merge_udtf_through_session():
    """TEST:2062-2404 through the host layer: argument grammar, output schema, contig order, strictness, min_dist"""
    import pyarrow as pa

    import bio_ranges_b200 as brg
    from bio_ranges_b200.session import NOT_IMPLEMENTED, PLAN, create_bio_session

    def tbl(rows, names=("contig", "pos_start", "pos_end"), t=pa.int64()):
        return pa.table({names[0]: pa.array([r[0] for r in rows], pa.string()), names[1]: pa.array([r[1] for r in rows], t),
                         names[2]: pa.array([r[2] for r in rows], t)})

    def rows_of(batches):
        t = pa.Table.from_batches(batches) if batches else None
        return [] if t is None else list(zip(*[t.column(i).to_pylist() for i in range(4)]))

    ctx = create_bio_session()
    for rows, md, strict, want in LITERALS:
        ctx.register_table("intervals", tbl(rows, t=pa.int32()))
        args = ("contig", "pos_start", "pos_end", "strict") if strict else ()
        assert rows_of(ctx.merge("intervals", *args, min_dist=md)) == want
    out = ctx.merge("intervals")
    t = pa.Table.from_batches(out)
    assert t.schema.names == ["contig", "pos_start", "pos_end", "n_intervals"]
    assert [str(f.type) for f in t.schema] == ["string", "int64", "int64", "int64"] and not any(f.nullable for f in t.schema)
    # TEST:2285-2317 custom column names come from the arguments
    ctx.register_table("c", tbl([("a", 100, 200), ("a", 150, 250)], names=("chr", "s", "e"), t=pa.int32()))
    t = pa.Table.from_batches(ctx.merge("c", "chr", "s", "e", min_dist=0))
    assert t.schema.names == ["chr", "s", "e", "n_intervals"] and rows_of([t.to_batches()[0]]) == [("a", 100, 250, 2)]
    # batches of the session batch size
    rng = np.random.default_rng(0)
    st = np.sort(rng.integers(0, 10**9, 50_000)) * 2
    ctx.register_table("big", pa.table({"contig": pa.array(["x"] * len(st)), "pos_start": pa.array(st, pa.int64()), "pos_end": pa.array(st, pa.int64())}))
    out = ctx.merge("big")
    assert [b.num_rows for b in out][:2] == [8192, 8192] and sum(b.num_rows for b in out) == len(np.unique(st))
    # errors
    with pytest.raises(brg.BioRangesError) as ei:
        ctx.merge("intervals", min_dist=-1)
    assert ei.value.status == PLAN and "merge() min_dist must be >= 0, got -1" in str(ei.value)
    with pytest.raises(brg.BioRangesError) as ei:
        ctx.merge("intervals", "a", "b")
    assert ei.value.status == PLAN and "merge() expects 0 or 3 column name arguments (got 2)" in str(ei.value)
    ctx.register_table("wide", tbl([("a", 1, 2**40)]))
    with pytest.raises(brg.BioRangesError) as ei:
        ctx.merge("wide")
    assert ei.value.status == NOT_IMPLEMENTED
    ctx.close()


# ---- cluster (tests/integration_test.rs:2410-2760, 3594-3626) ------------------------------------------------------------
CLUSTER_LITERALS = [  # (rows, min_dist, strict, expected rows sorted by (contig, start, end))
    ([("a", 100, 200), ("a", 150, 250), ("a", 400, 500)], 0, False,
     [("a", 100, 200, 0, 100, 250), ("a", 150, 250, 0, 100, 250), ("a", 400, 500, 1, 400, 500)]),
    ([("a", 100, 200), ("a", 210, 300)], 0, False, [("a", 100, 200, 0, 100, 200), ("a", 210, 300, 1, 210, 300)]),
    ([("a", 100, 200), ("a", 210, 300)], 10, False, [("a", 100, 200, 0, 100, 300), ("a", 210, 300, 0, 100, 300)]),
    ([("a", 100, 200), ("a", 150, 250), ("b", 100, 200), ("b", 300, 400)], 0, False,
     [("a", 100, 200, 0, 100, 250), ("a", 150, 250, 0, 100, 250), ("b", 100, 200, 1, 100, 200), ("b", 300, 400, 2, 300, 400)]),
    ([("a", 100, 200)], 0, False, [("a", 100, 200, 0, 100, 200)]),
    ([], 0, False, []),
    ([("a", 100, 150), ("a", 150, 200)], 0, False, [("a", 100, 150, 0, 100, 200), ("a", 150, 200, 0, 100, 200)]),
    ([("a", 100, 150), ("a", 150, 200)], 0, True, [("a", 100, 150, 0, 100, 150), ("a", 150, 200, 1, 150, 200)]),
    (RANGES_READS, 0, False, [
        ("chr1", 150, 250, 0, 150, 700), ("chr1", 190, 300, 0, 150, 700), ("chr1", 300, 501, 0, 150, 700), ("chr1", 500, 700, 0, 150, 700),
        ("chr1", 15000, 15000, 1, 15000, 15000), ("chr1", 22000, 22300, 2, 22000, 22300),
        ("chr2", 150, 250, 3, 150, 700), ("chr2", 190, 300, 3, 150, 700), ("chr2", 300, 500, 3, 150, 700), ("chr2", 500, 700, 3, 150, 700),
        ("chr2", 15000, 15000, 4, 15000, 15000), ("chr2", 22000, 22300, 5, 22000, 22300), ("chr3", 234, 300, 6, 234, 300)]),
]


@pytest.mark.parametrize("case", range(len(CLUSTER_LITERALS)))
def test_oracle_cluster_reference_literals(case):
    from oracle import oracle as O

    rows, min_dist, strict, want = CLUSTER_LITERALS[case]
    names, k, s, e = _encode_sorted(rows)
    row, cid, cs, ce = O.cluster(k, s, e, max(1, len(names)), min_dist, strict)
    got = [(rows[int(r)][0], rows[int(r)][1], rows[int(r)][2], int(c), int(a), int(b)) for r, c, a, b in zip(row, cid, cs, ce)]
    assert got == want


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_gpu_cluster_vs_oracle(seed):
    import bio_ranges_b200 as brg
    from oracle import oracle as O

    nk, k, s, e, md, strict = _random_case(seed)
    ctx = brg.default_context(0)
    grow, gcs, gce, gcid = ctx.cluster(k, s.astype(np.int32), e.astype(np.int32), nk, md, strict)
    orow, ocid, ocs, oce = O.cluster(k, s, e, nk, md, strict)
    assert (grow == orow).all() and (gcid == ocid).all() and (gcs == ocs).all() and (gce == oce).all()


@pytest.mark.gpu
def test_cluster_udtf_through_session():
    import pyarrow as pa

    import bio_ranges_b200 as brg
    from bio_ranges_b200.session import NOT_IMPLEMENTED, create_bio_session

    ctx = create_bio_session()
    for rows, md, strict, want in CLUSTER_LITERALS:
        ctx.register_table("intervals", pa.table({"contig": pa.array([r[0] for r in rows], pa.string()),
                                                  "pos_start": pa.array([r[1] for r in rows], pa.int32()),
                                                  "pos_end": pa.array([r[2] for r in rows], pa.int32())}))
        args = ("contig", "pos_start", "pos_end", "strict") if strict else ()
        out = ctx.cluster("intervals", *args, min_dist=md)
        got = [] if not out else list(zip(*[pa.Table.from_batches(out).column(i).to_pylist() for i in range(6)]))
        assert got == want
    t = pa.Table.from_batches(ctx.cluster("intervals"))
    assert t.schema.names == ["contig", "pos_start", "pos_end", "cluster", "cluster_start", "cluster_end"]
    assert [str(f.type) for f in t.schema] == ["string"] + ["int64"] * 5  # a 3-column input is rebuilt as Utf8 / Int64
    # TEST:3594-3626 extra columns are preserved with their types
    ctx.register_table("x", pa.table({"contig": ["a", "a", "a"], "pos_start": pa.array([400, 100, 150], pa.int32()),
                                      "pos_end": pa.array([500, 200, 250], pa.int32()), "gene": ["TP53", "BRCA1", "BRCA2"],
                                      "score": pa.array([0.75, 0.95, 0.85], pa.float64())}))
    t = pa.Table.from_batches(ctx.cluster("x"))
    assert t.schema.names == ["contig", "pos_start", "pos_end", "gene", "score", "cluster", "cluster_start", "cluster_end"]
    assert t.schema.field("pos_start").type == pa.int32()
    assert list(zip(*[t.column(i).to_pylist() for i in range(8)])) == [
        ("a", 100, 200, "BRCA1", 0.95, 0, 100, 250), ("a", 150, 250, "BRCA2", 0.85, 0, 100, 250), ("a", 400, 500, "TP53", 0.75, 1, 400, 500)]
    # TEST:2638-2688 needs 64-bit coordinates: declared unsupported on the 32-bit GPU path
    ctx.register_table("w", pa.table({"contig": ["a", "a"], "pos_start": pa.array([2**63 - 8, 2**63 - 1], pa.int64()),
                                      "pos_end": pa.array([2**63 - 2, 2**63 - 1], pa.int64())}))
    with pytest.raises(brg.BioRangesError) as ei:
        ctx.cluster("w", min_dist=100)
    assert ei.value.status == NOT_IMPLEMENTED
    ctx.close()


# ---- subtract (tests/integration_test.rs:3185-3712) ----------------------------------------------------------------------
SUBTRACT_LITERALS = [  # (left rows, right rows, strict, expected pieces)
    ([("a", 100, 400)], [("a", 200, 300)], False, [("a", 100, 200), ("a", 300, 400)]),
    ([("a", 100, 200)], [("a", 300, 400)], False, [("a", 100, 200)]),
    ([("a", 100, 200)], [("a", 50, 300)], False, []),
    ([("a", 100, 500)], [("a", 150, 200), ("a", 300, 350)], False, [("a", 100, 150), ("a", 200, 300), ("a", 350, 500)]),
    ([("a", 100, 300), ("b", 100, 300)], [("a", 150, 250)], False, [("a", 100, 150), ("a", 250, 300), ("b", 100, 300)]),
    ([], [("a", 1, 2)], False, []),
    ([("a", 100, 200)], [], False, [("a", 100, 200)]),
    ([("a", 100, 300)], [("a", 300, 400)], False, [("a", 100, 300)]),
    ([("a", 100, 300)], [("a", 300, 400)], True, [("a", 100, 300)]),
    ([("a", 100, 500)], [("a", 150, 250), ("a", 200, 350)], False, [("a", 100, 150), ("a", 350, 500)]),
]


def _encode_two(left, right):
    names = sorted({r[0] for r in left})
    ids = {n: i for i, n in enumerate(names)}
    enc = lambda rows: (np.array([ids.get(r[0], 0xFFFFFFFF) for r in rows], np.uint32), np.array([r[1] for r in rows], np.int64),
                        np.array([r[2] for r in rows], np.int64))
    return names, enc(left), enc(right)


@pytest.mark.parametrize("case", range(len(SUBTRACT_LITERALS)))
def test_oracle_subtract_reference_literals(case):
    from oracle import oracle as O

    left, right, strict, want = SUBTRACT_LITERALS[case]
    names, (lk, ls, le), (rk, rs, re) = _encode_two(left, right)
    row, ps, pe = O.subtract(lk, ls, le, rk, rs, re, max(1, len(names)), strict)
    assert [(left[int(r)][0], int(a), int(b)) for r, a, b in zip(row, ps, pe)] == want


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_gpu_subtract_vs_oracle(seed):
    import bio_ranges_b200 as brg
    from oracle import oracle as O

    rng = np.random.default_rng(100 + seed)
    nk = int(rng.choice([1, 4, 40]))
    nl, nr = int(rng.choice([1, 300, 50_000])), int(rng.choice([0, 1, 500, 80_000]))
    span = int(rng.choice([2000, 300_000, 1_000_000_000]))
    mk = lambda n, maxlen: (rng.integers(0, nk + (1 if seed % 4 == 0 else 0), n).astype(np.uint32), rng.integers(0, span, n).astype(np.int64))
    lk, ls = mk(nl, 0)
    lk = np.minimum(lk, nk - 1)
    le = ls + rng.integers(0, int(rng.choice([1, 100, 20000])), nl)
    rk, rs = mk(nr, 0)
    re = rs + rng.integers(0, int(rng.choice([1, 60, 9000])), nr)
    strict = bool(seed % 2)
    ctx = brg.default_context(0)
    gi = ctx.build(rk, rs.astype(np.int32), re.astype(np.int32), nk + 1, brg.WANT_COVERAGE)
    order = ctx.sort_rows(lk, ls.astype(np.int32), le.astype(np.int32), nk)
    idx, ps, pe = gi.subtract(lk[order], ls[order].astype(np.int32), le[order].astype(np.int32), strict=strict)
    orow, ops, ope = O.subtract(lk, ls, le, np.where(rk < nk, rk, 0xFFFFFFFF).astype(np.uint32), rs, re, nk, strict)
    assert len(idx) == len(orow)
    assert (order[idx] == orow).all() and (ps == ops).all() and (pe == ope).all()
    gi.close()


@pytest.mark.gpu
def test_subtract_udtf_through_session():
    import pyarrow as pa

    from bio_ranges_b200.session import create_bio_session

    def tbl(rows, names=("contig", "pos_start", "pos_end")):
        return pa.table({names[0]: pa.array([r[0] for r in rows], pa.string()), names[1]: pa.array([r[1] for r in rows], pa.int32()),
                         names[2]: pa.array([r[2] for r in rows], pa.int32())})

    ctx = create_bio_session()
    for left, right, strict, want in SUBTRACT_LITERALS:
        ctx.register_table("left_t", tbl(left))
        ctx.register_table("right_t", tbl(right))
        args = ("contig", "pos_start", "pos_end", "strict") if strict else ()
        out = ctx.subtract("left_t", "right_t", *args)
        got = [] if not out else list(zip(*[pa.Table.from_batches(out).column(i).to_pylist() for i in range(3)]))
        assert got == want
        if out:
            assert [str(f.type) for f in out[0].schema] == ["string", "int64", "int64"]
    # TEST:3628-3712 extra columns preserved, start / end replaced by the Int64 pieces
    ctx.register_table("left_multi", pa.table({"contig": ["a"], "pos_start": pa.array([100], pa.int32()), "pos_end": pa.array([600], pa.int32()),
                                               "gene": ["BRCA1"], "score": pa.array([0.95], pa.float64())}))
    ctx.register_table("right_multi", tbl([("a", 200, 300), ("a", 400, 500)]))
    t = pa.Table.from_batches(ctx.subtract("left_multi", "right_multi"))
    assert t.schema.names == ["contig", "pos_start", "pos_end", "gene", "score"]
    assert t.schema.field("pos_start").type == pa.int64() and t.schema.field("score").type == pa.float64()
    assert list(zip(*[t.column(i).to_pylist() for i in range(5)])) == [
        ("a", 100, 200, "BRCA1", 0.95), ("a", 300, 400, "BRCA1", 0.95), ("a", 500, 600, "BRCA1", 0.95)]
    # TEST:3453-3495 separate column names per side
    ctx.register_table("lc", tbl([("a", 100, 400)], names=("chr", "s", "e")))
    ctx.register_table("rc", tbl([("a", 200, 300)], names=("c2", "s2", "e2")))
    t = pa.Table.from_batches(ctx.subtract("lc", "rc", "chr", "s", "e", "c2", "s2", "e2"))
    assert t.schema.names == ["chr", "s", "e"] and t.column("s").to_pylist() == [100, 300]
    ctx.close()


# ---- complement (tests/integration_test.rs:2821-3180, 3803-3882) -----------------------------------------------------------
I64MAX = 2**63 - 1
COMPLEMENT_LITERALS = [  # (input rows, view rows or None, strict, expected gaps)
    ([("a", 100, 200), ("a", 300, 400)], None, False, [("a", 0, 100), ("a", 200, 300), ("a", 400, I64MAX)]),
    ([("a", 100, 200), ("b", 300, 400)], None, False, [("a", 0, 100), ("a", 200, I64MAX), ("b", 0, 300), ("b", 400, I64MAX)]),
    ([("a", 100, 200), ("a", 300, 400)], [("a", 0, 500)], False, [("a", 0, 100), ("a", 200, 300), ("a", 400, 500)]),
    ([("a", 0, 500)], [("a", 0, 500)], False, []),
    ([], [("a", 0, 500)], False, [("a", 0, 500)]),
    ([("a", 100, 200), ("b", 300, 400)], [("a", 0, 500), ("b", 0, 600)], False, [("a", 0, 100), ("a", 200, 500), ("b", 0, 300), ("b", 400, 600)]),
    ([("a", 100, 250), ("a", 200, 400)], [("a", 0, 500)], False, [("a", 0, 100), ("a", 400, 500)]),
    ([("a", 100, 200)], [("a", 0, 500), ("b", 0, 300)], False, [("a", 0, 100), ("a", 200, 500), ("b", 0, 300)]),
    # the partitioned-parquet test (TEST:3760-3882): two input files, one view file
    ([("chr1", 0, 10), ("chr1", 20, 30), ("chr2", 10, 20), ("chr1", 8, 25), ("chr2", 30, 40)], [("chr1", 0, 40), ("chr2", 0, 50)], False,
     [("chr1", 30, 40), ("chr2", 0, 10), ("chr2", 20, 30), ("chr2", 40, 50)]),
    ([("chr1", 0, 10), ("chr1", 20, 30), ("chr2", 10, 20), ("chr1", 8, 25), ("chr2", 30, 40)], None, False,
     [("chr1", 30, I64MAX), ("chr2", 0, 10), ("chr2", 20, 30), ("chr2", 40, I64MAX)]),
]


def _encode_complement(rows, views):
    """contig ids in name order; views in emission order: input contigs by name with their views sorted by (start, end) or
    the implicit (0, i64::MAX), then view-only contigs by name"""
    names = sorted({r[0] for r in rows})
    ids = {n: i for i, n in enumerate(names)}
    k = np.array([ids[r[0]] for r in rows], np.uint32)
    s, e = np.array([r[1] for r in rows], np.int64), np.array([r[2] for r in rows], np.int64)
    vrows = []
    for n in names:
        mine = sorted((v[1], v[2]) for v in (views or []) if v[0] == n)
        vrows += [(ids[n], n, a, b) for a, b in mine] if mine else [(ids[n], n, 0, I64MAX)]
    for n in sorted({v[0] for v in (views or [])} - set(names)):
        vrows += [(0xFFFFFFFF, n, a, b) for a, b in sorted((v[1], v[2]) for v in views if v[0] == n)]
    vk = np.array([v[0] for v in vrows], np.uint32)
    vs, ve = np.array([v[2] for v in vrows], np.int64), np.array([v[3] for v in vrows], np.int64)
    return names, (k, s, e), [v[1] for v in vrows], (vk, vs, ve)


@pytest.mark.parametrize("case", range(len(COMPLEMENT_LITERALS)))
def test_oracle_complement_reference_literals(case):
    from oracle import oracle as O

    rows, views, strict, want = COMPLEMENT_LITERALS[case]
    names, (k, s, e), vcontig, (vk, vs, ve) = _encode_complement(rows, views)
    ov, gs, ge = O.complement(k, s, e, max(1, len(names)), vk, vs, ve, strict)
    assert [(vcontig[int(v)], int(a), int(b)) for v, a, b in zip(ov, gs, ge)] == want


def test_oracle_complement_strict_keeps_touching_runs_apart():
    from oracle import oracle as O

    # strict: s < cur_end merges, so [100,150] and [150,200] stay two runs; no gap between them either way
    k, s, e = np.zeros(2, np.uint32), np.array([100, 150], np.int64), np.array([150, 200], np.int64)
    for strict in (False, True):
        ov, gs, ge = O.complement(k, s, e, 1, np.zeros(1, np.uint32), np.array([0], np.int64), np.array([300], np.int64), strict)
        assert list(zip(gs.tolist(), ge.tolist())) == [(0, 100), (200, 300)]
    # a zero-width run splits the gap it sits in (interval_start > cursor, cursor = interval_end)
    ov, gs, ge = O.complement(np.zeros(1, np.uint32), [50], [50], 1, np.zeros(1, np.uint32), [0], [100], False)
    assert list(zip(gs.tolist(), ge.tolist())) == [(0, 50), (50, 100)]
    # a degenerate view of an input contig yields nothing; of a view-only contig it comes out as is
    ov, gs, ge = O.complement(np.zeros(1, np.uint32), [50], [60], 1, np.array([0, 0xFFFFFFFF], np.uint32), [70, 70], [20, 20], False)
    assert list(zip(ov.tolist(), gs.tolist(), ge.tolist())) == [(1, 70, 20)]


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(14))
def test_gpu_complement_vs_oracle(seed):
    import bio_ranges_b200 as brg
    from oracle import oracle as O

    rng = np.random.default_rng(300 + seed)
    nk = int(rng.choice([1, 4, 40]))
    n = int(rng.choice([0, 1, 300, 50_000, 400_000]))
    nv = int(rng.choice([1, 5, 200, 20_000]))
    span = int(rng.choice([2000, 300_000, 1_000_000_000]))
    k = rng.integers(0, nk, n).astype(np.uint32)
    s = rng.integers(0, span, n).astype(np.int64)
    e = s + rng.integers(0, int(rng.choice([1, 100, 20000])), n)
    # views: some keys absent (>= nk or the ABSENT id), some degenerate, some implicit whole-contig, overlapping each other
    vk = rng.integers(0, nk + 2, nv).astype(np.uint32)
    vk[vk == nk + 1] = 0xFFFFFFFF
    vs = rng.integers(-50, span, nv).astype(np.int64)
    ve = vs + rng.integers(-5, int(rng.choice([10, 5000, span])), nv)
    whole = rng.random(nv) < 0.1
    vs[whole], ve[whole] = 0, I64MAX
    order = np.lexsort((ve, vs, vk))
    vk, vs, ve = vk[order], vs[order], ve[order]
    strict = bool(seed % 2)
    ctx = brg.default_context(0)
    gv, gs, ge = ctx.complement(k, s.astype(np.int32), e.astype(np.int32), nk, vk, vs, ve, strict=strict)
    ov, os_, oe = O.complement(k, s, e, nk, vk, vs, ve, strict)
    assert len(gv) == len(ov)
    assert (gv == ov).all() and (gs == os_).all() and (ge == oe).all()


@pytest.mark.gpu
def test_complement_udtf_through_session():
    import pyarrow as pa

    from bio_ranges_b200.session import create_bio_session

    def tbl(rows, names=("contig", "pos_start", "pos_end")):
        return pa.table({names[0]: pa.array([r[0] for r in rows], pa.string()), names[1]: pa.array([r[1] for r in rows], pa.int32()),
                         names[2]: pa.array([r[2] for r in rows], pa.int32())})

    ctx = create_bio_session()
    for rows, views, strict, want in COMPLEMENT_LITERALS:
        ctx.register_table("intervals", tbl(rows))
        args = []
        if views is not None:
            ctx.register_table("chromsizes", tbl(views))
            args.append("chromsizes")
        if strict:
            args.append("strict")
        out = ctx.complement("intervals", *args)
        got = [] if not out else list(zip(*[pa.Table.from_batches(out).column(i).to_pylist() for i in range(3)]))
        assert got == want  # already in the reference's emission order (contig name, view, coordinate)
        if out:
            assert [str(f.type) for f in out[0].schema] == ["string", "int64", "int64"]
            assert not any(f.nullable for f in out[0].schema)
    # TEST:3093-3135 custom column names, used for both tables; the output takes the input's names
    ctx.register_table("iv", tbl([("a", 100, 200)], names=("chr", "s", "e")))
    ctx.register_table("cs", tbl([("a", 0, 500)], names=("chr", "s", "e")))
    t = pa.Table.from_batches(ctx.complement("iv", "cs", "chr", "s", "e"))
    assert t.schema.names == ["chr", "s", "e"] and list(zip(t.column("s").to_pylist(), t.column("e").to_pylist())) == [(0, 100), (200, 500)]
    # separate view column names (table_function.rs:748-766)
    ctx.register_table("cs2", tbl([("a", 0, 500)], names=("vc", "vs", "ve")))
    t = pa.Table.from_batches(ctx.complement("iv", "cs2", "chr", "s", "e", "vc", "vs", "ve"))
    assert t.schema.names == ["chr", "s", "e"] and t.num_rows == 2
    # argument-count error text (table_function.rs:757-763)
    with pytest.raises(Exception, match=r"complement\(\) unexpected number of arguments \(2 extra args after table name\)"):
        ctx.complement("iv", "cs", "chr")
    ctx.close()
