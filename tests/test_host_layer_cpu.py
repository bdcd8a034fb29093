"""Host layer (include/bioranges_host.h) pieces that need no GPU: BioConfig / Algorithm parsing, the table functions'
argument grammar and errors, column resolution (PosArray::resolve) and `take`.  Texts are the reference's
(session_context.rs, table_function.rs, array_utils.rs)."""
import os
import re

import numpy as np
import pyarrow as pa
import pytest

import bio_ranges_b200 as brg
from bio_ranges_b200 import session as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _err(fn, *a, **kw):
    with pytest.raises(brg.BioRangesError) as ei:
        fn(*a, **kw)
    return ei.value.status, str(ei.value)


def test_host_header_and_binding_agree():
    src = open(os.path.join(ROOT, "include", "bioranges_host.h")).read()
    declared = sorted(set(re.findall(r"BRG_API[^;(]*?\b(brgh_\w+)\s*\(", src)))
    assert declared == sorted(S.HOST_SYMBOLS)
    L = brg._cabi.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_bio_config_set_show():
    """session_context.rs:48-110 -- defaults, case-insensitive Algorithm parsing, Display names, parse error text"""
    ctx = S.create_bio_session()
    assert ctx.get("bio.interval_join_algorithm") == "Coitrees"
    assert ctx.get("bio.prefer_interval_join") == "true"
    assert ctx.get("bio.interval_join_low_memory") == "false"
    for text, shown in [("lapper", "Lapper"), ("INTERVALTREE", "IntervalTree"), ("ArrayIntervalTree", "ArrayIntervalTree"),
                        ("superintervals", "SuperIntervals"), ("CoitreesNearest", "CoitreesNearest"),
                        ("coitreescountoverlaps", "CoitreesCountOverlaps"), ("Coitrees", "Coitrees")]:
        ctx.set("bio.interval_join_algorithm", text)
        assert ctx.get("bio.interval_join_algorithm") == shown
        assert ctx.explain_interval_join().endswith(f"alg={shown}")  # EXPLAIN text the reference's tests grep (TEST:129,223)
    st, msg = _err(ctx.set, "bio.interval_join_algorithm", "btree")
    assert st == S.EXTERNAL and "Can't parse 'btree' as Algorithm" in msg
    ctx.set("bio.interval_join_low_memory", "true")
    assert ctx.get("bio.interval_join_low_memory") == "true"
    assert _err(ctx.set, "bio.no_such_option", "1")[0] == S.EXTERNAL
    ctx.close()


def _tbl(rows, names=("contig", "pos_start", "pos_end"), pos_type=pa.int32()):
    return pa.table({names[0]: pa.array([r[0] for r in rows], pa.string()),
                     names[1]: pa.array([r[1] for r in rows], pos_type),
                     names[2]: pa.array([r[2] for r in rows], pos_type)})


def test_udtf_argument_grammar_errors():
    """table_function.rs:59-214 -- errors are raised before any device work"""
    ctx = S.create_bio_session()
    ctx.register_table("a", _tbl([("chr1", 1, 5)]))
    ctx.register_table("b", _tbl([("chr1", 2, 3)]))
    st, msg = _err(ctx.count_overlaps, "a", "b", "contig", "pos_start")
    assert st == S.PLAN and "count_overlaps() expects 0, 3, or 6 column name arguments (got 2)" in msg
    assert "Usage: count_overlaps('left_table', 'right_table' [, col1, col2, col3 [, col4, col5, col6]] [, 'strict'|'weak'])" in msg
    st, msg = _err(ctx.coverage, "a", "b", "c`x", "s", "e")
    assert st == S.PLAN and "coverage() column name must not contain backtick characters, got: c`x" in msg
    st, msg = _err(ctx.nearest, "a", "b", k=0)
    assert st == S.PLAN and "nearest() k must be >= 1" in msg
    st, msg = _err(ctx.count_overlaps, "a", "nope")
    assert st == S.PLAN and "table 'nope' not found" in msg
    # column errors (array_utils.rs:178-300)
    st, msg = _err(ctx.count_overlaps, "a", "b", "chrom", "pos_start", "pos_end")
    assert st == S.PLAN and "contig column 'chrom' not found in batch with columns: [\"contig\", \"pos_start\", \"pos_end\"]" in msg
    ctx.register_table("f", pa.table({"contig": ["chr1"], "pos_start": pa.array([1.5], pa.float64()), "pos_end": pa.array([2], pa.int32())}))
    st, msg = _err(ctx.count_overlaps, "f", "b")
    assert st == S.NOT_IMPLEMENTED
    assert "unsupported data type Float64 for start column 'pos_start'; expected Int32, Int64, UInt32, or UInt64" in msg
    ctx.register_table("g", pa.table({"contig": pa.array([7], pa.int32()), "pos_start": pa.array([1], pa.int32()), "pos_end": pa.array([2], pa.int32())}))
    st, msg = _err(ctx.count_overlaps, "g", "b")
    assert st == S.NOT_IMPLEMENTED and "unsupported data type Int32 for contig column 'contig'; expected Utf8, LargeUtf8, or Utf8View" in msg
    ctx.close()


def test_sweep_function_argument_errors():
    """merge / cluster / complement argument grammar (table_function.rs:236-241, 472-520, 615-767): raised before any device work"""
    ctx = S.create_bio_session()
    ctx.register_table("a", _tbl([("chr1", 1, 5)]))
    st, msg = _err(ctx.merge, "a", min_dist=-1)
    assert st == S.PLAN and "merge() min_dist must be >= 0, got -1" in msg
    st, msg = _err(ctx.cluster, "a", "x", "y")
    assert st == S.PLAN and "cluster() expects 0 or 3 column name arguments (got 2)" in msg
    st, msg = _err(ctx.complement, "a", "views", "c")
    assert st == S.PLAN and "complement() unexpected number of arguments (2 extra args after table name)" in msg
    assert "Usage: complement('table' [, 'view_table'] [, col1, col2, col3 [, vcol1, vcol2, vcol3]] [, 'strict'|'weak'])" in msg
    st, msg = _err(ctx.complement, "a", "views", "c", "s", "e", "strict", "x", "y", "z", "w")
    assert st == S.PLAN and "9 extra args" in msg  # (a trailing 'strict' is only recognised in last position)
    st, msg = _err(ctx.complement, "nope")
    assert st == S.PLAN and "table 'nope' not found" in msg
    st, msg = _err(ctx.complement, "a", "no_view")
    assert st == S.PLAN and "table 'no_view' not found" in msg
    ctx.close()


def test_resolve_positions_like_posarray_resolve():
    """array_utils.rs:73-131"""
    b = pa.record_batch({"i32": pa.array([1, -5, 2**31 - 1], pa.int32()), "i64": pa.array([1, 2, 3], pa.int64()),
                         "u32": pa.array([0, 7, 2**31 - 1], pa.uint32()), "u64": pa.array([5, 6, 7], pa.uint64()),
                         "big": pa.array([1, 2**31, 3], pa.int64()), "ubig": pa.array([1, 2, 2**32 + 5], pa.uint64()),
                         "nul": pa.array([1, None, 3], pa.int64()), "txt": pa.array(["a", "b", "c"])})
    assert list(S.resolve_positions(b, "i32")) == [1, -5, 2**31 - 1]
    assert list(S.resolve_positions(b, "i64")) == [1, 2, 3]
    assert list(S.resolve_positions(b, "u32")) == [0, 7, 2**31 - 1]
    assert list(S.resolve_positions(b, "u64")) == [5, 6, 7]
    st, msg = _err(S.resolve_positions, b, "big")
    assert st == S.EXECUTION and "coordinate value 2147483648 at row 1 overflows i32 (max 2147483647)" in msg
    st, msg = _err(S.resolve_positions, b, "ubig")
    assert st == S.EXECUTION and "coordinate value 4294967301 at row 2 overflows i32 (max 2147483647)" in msg
    st, msg = _err(S.resolve_positions, b, "nul")
    assert st == S.EXECUTION and "coordinate column contains null values" in msg
    st, msg = _err(S.resolve_positions, b, "txt", "end")
    assert st == S.NOT_IMPLEMENTED and "unsupported data type Utf8 for end column 'txt'" in msg
    # a sliced batch (non-zero offset)
    assert list(S.resolve_positions(b.slice(1, 2), "i64")) == [2, 3]


def test_take_matches_arrow_compute():
    import pyarrow.compute as pc

    rng = np.random.default_rng(0)
    n = 1000
    b = pa.record_batch({
        "s": pa.array([None if i % 17 == 0 else f"chr{i % 23}_{'x' * (i % 5)}" for i in range(n)], pa.string()),
        "ls": pa.array([f"L{i}" for i in range(n)], pa.large_string()),
        "i32": pa.array(rng.integers(-100, 100, n), pa.int32()),
        "i64": pa.array([None if i % 11 == 0 else int(i) * 10**10 for i in range(n)], pa.int64()),
        "f64": pa.array(rng.random(n), pa.float64()),
        "u8": pa.array(rng.integers(0, 255, n), pa.uint8()),
        "flag": pa.array([bool(i % 3) for i in range(n)], pa.bool_()),
    })
    idx = rng.integers(0, n, 5000).astype(np.uint32)
    valid = rng.random(5000) > 0.2
    for src in (b, b.slice(100, 700)):
        ii = idx % src.num_rows
        got = S.take(src, ii, valid, prefix="left_")
        want_idx = pa.array([int(v) if ok else None for v, ok in zip(ii, valid)], pa.uint32())
        for name in src.schema.names:
            assert got.column("left_" + name).equals(pc.take(src.column(name), want_idx)), name
        got2 = S.take(src, ii)
        assert got2.column("s").equals(pc.take(src.column("s"), pa.array(ii, pa.uint32())))
    st, msg = _err(S.take, b, np.array([n], np.uint32))
    assert st == S.ARROW and "index out of bounds" in msg.lower()


# ---- the range-join filter parser: physical_planner/intervals.rs:220-525, the reference's own test cases -----------------------
# tables of the reference test: a(contig, l_start, l_end) is the SQL-left input, b(contig, name, r_end, r_start) the right one
GTEQ_LTEQ = ["b.r_end >= a.l_start AND a.l_end >= b.r_start", "b.r_end >= a.l_start AND b.r_start <= a.l_end",
             "a.l_start <= b.r_end AND a.l_end >= b.r_start", "a.l_start <= b.r_end AND b.r_start <= a.l_end",
             "a.l_end >= b.r_start AND b.r_end >= a.l_start", "a.l_end >= b.r_start AND a.l_start <= b.r_end",
             "b.r_start <= a.l_end AND b.r_end >= a.l_start", "b.r_start <= a.l_end AND a.l_start <= b.r_end"]
GT_LT = ["b.r_end > a.l_start AND a.l_end > b.r_start", "b.r_end > a.l_start AND b.r_start < a.l_end",
         "a.l_start < b.r_end AND a.l_end > b.r_start", "a.l_start < b.r_end AND b.r_start < a.l_end",
         "a.l_end > b.r_start AND b.r_end > a.l_start", "a.l_end > b.r_start AND a.l_start < b.r_end",
         "b.r_start < a.l_end AND b.r_end > a.l_start", "b.r_start < a.l_end AND a.l_start < b.r_end"]


@pytest.mark.parametrize("cond", GTEQ_LTEQ)
def test_join_filter_all_weak_orderings(cond):
    """intervals.rs:246-338 test_all_comp_combinations_for_gteq_lteq"""
    got = S.parse_join_filter(cond, "a", "b")
    assert got == {"left_start": "l_start", "left_end": "l_end", "right_start": "r_start", "right_end": "r_end",
                   "left_end_minus_one": False, "right_end_minus_one": False}


@pytest.mark.parametrize("cond", GT_LT)
def test_join_filter_all_strict_orderings(cond):
    """intervals.rs:347-462 test_all_comp_combinations_for_gt_lt: both ends become `end - 1`"""
    got = S.parse_join_filter(cond, "a", "b")
    assert got == {"left_start": "l_start", "left_end": "l_end", "right_start": "r_start", "right_end": "r_end",
                   "left_end_minus_one": True, "right_end_minus_one": True}


def test_join_filter_mixed_strictness_or_and_complex():
    # intervals.rs:464-490: only the strict comparison's END gets the minus one
    got = S.parse_join_filter("b.r_start < a.l_end AND a.l_start <= b.r_end", "a", "b")
    assert got["left_end_minus_one"] and not got["right_end_minus_one"]
    got = S.parse_join_filter("b.r_start <= a.l_end AND a.l_start < b.r_end", "a", "b")
    assert not got["left_end_minus_one"] and got["right_end_minus_one"]
    # intervals.rs:340-342: OR is not a range join
    st, msg = _err(S.parse_join_filter, "b.r_start <= a.l_end OR a.l_start <= b.r_end", "a", "b")
    assert st == S.INTERNAL and "filter expression is not an AND" in msg
    # intervals.rs:494-499: an operand over two columns ("complex sub queries are not supported": the reference panics)
    st, msg = _err(S.parse_join_filter, "(b.r_end - a.l_start) >= a.l_start AND a.l_end >= b.r_start", "a", "b")
    assert "complex sub queries are not supported" in msg
    # parse_condition's other refusals, with the reference's texts
    assert "Unsupported operator: =" in _err(S.parse_join_filter, "a.l_start = b.r_end AND a.l_end >= b.r_start", "a", "b")[1]
    assert "couldn't parse as ls </<= re" in _err(S.parse_join_filter, "a.l_start <= a.l_end AND a.l_end >= b.r_start", "a", "b")[1]
    assert "already set" in _err(S.parse_join_filter, "a.l_start <= b.r_end AND a.l_end <= b.r_start", "a", "b")[1]
    # the form the reference's own OverlapProvider writes (overlap.rs:171-181): CASTs, quoted names, case-insensitive aliases
    got = S.parse_join_filter("CAST(a.`pos_end` AS INTEGER) >= CAST(b.`pos_start` AS INTEGER) and CAST(A.\"pos_start\" AS INTEGER) <= CAST(B.pos_end AS INTEGER)", "a", "b")
    assert got["left_start"] == "pos_start" and got["left_end"] == "pos_end" and got["right_start"] == "pos_start" and got["right_end"] == "pos_end"
    got = S.parse_join_filter("(a.s <= b.e) AND (a.e > b.s)", "a", "b")
    assert got == {"left_start": "s", "left_end": "e", "right_start": "s", "right_end": "e", "left_end_minus_one": True, "right_end_minus_one": False}


def test_ingest_contig_runs_from_dictionary_column(tmp_path):
    """ingest.py: a dictionary-encoded contig column becomes key runs without touching a string per row"""
    import pyarrow.parquet as pq

    from bio_ranges_b200 import ingest

    names = ["chr1"] * 5 + ["chr2"] * 3 + ["chrUn"] * 2 + ["chr1"] * 1
    t = pa.table({"contig": pa.array(names), "pos_start": pa.array(range(11), pa.int64()), "pos_end": pa.array(range(1, 12), pa.int64())})
    pq.write_table(t, str(tmp_path / "t.parquet"))
    (rb,) = list(ingest.iter_parquet(str(tmp_path / "t.parquet"), ("contig", "pos_start", "pos_end")))
    assert pa.types.is_dictionary(rb.column(0).type)  # decoded from the dictionary pages, not re-encoded
    rk, rend = ingest.contig_runs(rb.column(0), {"chr1": 0, "chr2": 1})
    assert rk.tolist() == [0, 1, brg.KEY_ABSENT, 0] and rend.tolist() == [5, 8, 10, 11]
    assert ingest._column_values(rb.column(1), "start").dtype == np.int64
