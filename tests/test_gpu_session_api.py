"""The reference's integration tests (tests/integration_test.rs), replayed through the host layer of libbioranges.so:
Arrow record batches with STRING contigs in, record batches out -- session options, table functions and the range join
behave and name their columns as the reference does.  Expected tables are the reference's literals."""
import numpy as np
import pyarrow as pa
import pytest

from _fixtures import INTERVAL_READS, INTERVAL_TARGETS, RANGES_READS, RANGES_TARGETS, golden
from reference_cases import EXPECTED_EQUI

pytestmark = pytest.mark.gpu


def _tbl(rows, pos_type=pa.int64(), contig_type=pa.string()):
    """CSV-read tables have Utf8 contigs and Int64 positions in DataFusion"""
    return pa.table({"contig": pa.array([r[0] for r in rows], contig_type),
                     "pos_start": pa.array([r[1] for r in rows], pos_type),
                     "pos_end": pa.array([r[2] for r in rows], pos_type)})


@pytest.fixture()
def ctx():
    from bio_ranges_b200.session import create_bio_session

    c = create_bio_session()
    yield c
    c.close()


def _rows(batches, cols=None):
    t = pa.Table.from_batches(batches)
    cols = cols or t.schema.names
    return sorted(zip(*[t.column(c).to_pylist() for c in cols]), key=repr)


def _tab(batches, like):
    """table of the batches; an operator that emitted no batch at all yields the empty table"""
    return pa.Table.from_batches(batches) if batches else like.slice(0, 0)


def init_tables(ctx, **kw):  # testing/data/interval/{reads,targets}.csv
    ctx.register_table("reads", _tbl(INTERVAL_READS, **kw))
    ctx.register_table("targets", _tbl(INTERVAL_TARGETS, **kw))


def init_ranges_tables(ctx, **kw):  # testing/data/ranges/{reads,targets}.csv
    ctx.register_table("reads", _tbl(RANGES_READS, **kw))
    ctx.register_table("targets", _tbl(RANGES_TARGETS, **kw))


# ---- TEST:86-137 test_equi_and_range_condition, every algorithm setting --------------------------------------------------
@pytest.mark.parametrize("alg", ["Coitrees", "IntervalTree", "ArrayIntervalTree", "Lapper", "SuperIntervals"])
def test_equi_and_range_condition(ctx, alg):
    init_tables(ctx)
    ctx.set("bio.prefer_interval_join", "true")
    ctx.set("bio.interval_join_algorithm", alg)
    assert f"IntervalJoinExec: mode=CollectLeft, join_type=Inner" in ctx.explain_interval_join() and f"alg={alg}" in ctx.explain_interval_join()
    out = ctx.interval_join("reads", "targets", [("contig", "contig")], ("pos_start", "pos_end"), ("pos_start", "pos_end"))
    t = pa.Table.from_batches(out)
    assert t.schema.names == ["contig", "pos_start", "pos_end", "contig", "pos_start", "pos_end"]
    got = sorted(zip(*[t.column(i).to_pylist() for i in range(6)]))
    assert got == sorted(EXPECTED_EQUI)


# ---- TEST:139-231 test_range_condition: no equi key ----------------------------------------------------------------------
def test_range_condition(ctx):
    init_tables(ctx)
    out = ctx.interval_join("reads", "targets", [], ("pos_start", "pos_end"), ("pos_start", "pos_end"))
    t = pa.Table.from_batches(out)
    got = sorted(zip(*[t.column(i).to_pylist() for i in range(6)]))
    brute = sorted(a + b for a in INTERVAL_READS for b in INTERVAL_TARGETS if a[1] <= b[2] and a[2] >= b[1])
    assert len(got) == 32 and got == brute


# ---- TEST:233-369 test_all_gteq_lteq_conditions / test_all_gt_lt_conditions ----------------------------------------------
def test_all_weak_and_strict_conditions(ctx):
    b12 = [(11, 15), (10, 15), (10, 10), (9, 15), (5, 15), (4, 15), (4, 10), (6, 8), (4, 8), (4, 5), (5, 5), (4, 4)]
    ctx.register_table("a", pa.table({"start": pa.array([5], pa.int32()), "end": pa.array([10], pa.int32())}))
    ctx.register_table("b", pa.table({"start": pa.array([r[0] for r in b12], pa.int32()), "end": pa.array([r[1] for r in b12], pa.int32())}))
    weak = [(10, 15), (10, 10), (9, 15), (5, 15), (4, 15), (4, 10), (6, 8), (4, 8), (5, 5), (4, 5)]
    strict = [(9, 15), (5, 15), (4, 15), (4, 10), (6, 8), (4, 8)]
    for (l, r) in (("a", "b"), ("b", "a")):  # either table may be the SQL-left (build) side
        for lt_s, gt_s, want in ((False, False, weak), (True, True, strict)):
            out = ctx.interval_join(l, r, [], ("start", "end"), ("start", "end"), lt_strict=lt_s, gt_strict=gt_s)
            t = pa.Table.from_batches(out)
            bcols = (2, 3) if l == "a" else (0, 1)
            assert sorted(zip(t.column(bcols[0]).to_pylist(), t.column(bcols[1]).to_pylist())) == sorted(want)


# ---- the same conditions as SQL text through the filter parser (physical_planner/intervals.rs) -------------------------------
def test_range_join_from_filter_text(ctx):
    b12 = [(11, 15), (10, 15), (10, 10), (9, 15), (5, 15), (4, 15), (4, 10), (6, 8), (4, 8), (4, 5), (5, 5), (4, 4)]
    ctx.register_table("a", pa.table({"start": pa.array([5], pa.int32()), "end": pa.array([10], pa.int32())}))
    ctx.register_table("b", pa.table({"start": pa.array([r[0] for r in b12], pa.int32()), "end": pa.array([r[1] for r in b12], pa.int32())}))
    weak = [(10, 15), (10, 10), (9, 15), (5, 15), (4, 15), (4, 10), (6, 8), (4, 8), (5, 5), (4, 5)]
    strict = [(9, 15), (5, 15), (4, 15), (4, 10), (6, 8), (4, 8)]
    for cond, want in (("a.start <= b.end AND a.end >= b.start", weak), ("b.end >= a.start AND b.start <= a.end", weak),
                       ("a.start < b.end AND a.end > b.start", strict), ("b.start < a.end AND b.end > a.start", strict)):
        t = pa.Table.from_batches(ctx.interval_join_filter("a", "b", [], cond))
        assert sorted(zip(t.column(2).to_pylist(), t.column(3).to_pylist())) == sorted(want)
    from bio_ranges_b200 import BioRangesError

    with pytest.raises(BioRangesError, match="filter expression is not an AND"):
        ctx.interval_join_filter("a", "b", [], "a.start <= b.end OR a.end >= b.start")


# ---- TEST:371-418 test_nearest: CoitreesNearest join, two-column key, strict ------------------------------------------------
def test_join_coitrees_nearest_two_column_key(ctx):
    ctx.register_table("a", pa.table({"contig": ["a"], "strand": ["s"], "start": pa.array([5], pa.int32()), "end": pa.array([10], pa.int32())}))
    ctx.register_table("b", pa.table({"contig": ["a", "a", "a", "b"], "strand": ["s", "s", "x", "s"],
                                      "start": pa.array([11, 20, 0, 1], pa.int32()), "end": pa.array([13, 21, 1, 2], pa.int32())}))
    ctx.set("bio.interval_join_algorithm", "CoitreesNearest")
    out = ctx.interval_join("a", "b", [("contig", "contig"), ("strand", "strand")], ("start", "end"), ("start", "end"),
                            lt_strict=True, gt_strict=True)
    t = pa.Table.from_batches(out)
    rows = sorted(zip(*[t.column(i).to_pylist() for i in range(8)]), key=repr)
    want = [("a", "s", 5, 10, "a", "s", 11, 13), ("a", "s", 5, 10, "a", "s", 20, 21),
            (None, None, None, None, "a", "x", 0, 1), (None, None, None, None, "b", "s", 1, 2)]
    assert rows == sorted(want, key=repr)


# ---- interval_join.rs:1627-1636: CoitreesCountOverlaps = one row per probe row, an overlapping interval or NULL --------------
def test_join_coitrees_count_overlaps_mode(ctx):
    init_ranges_tables(ctx)
    ctx.set("bio.interval_join_algorithm", "CoitreesCountOverlaps")
    out = ctx.interval_join("reads", "targets", [("contig", "contig")], ("pos_start", "pos_end"), ("pos_start", "pos_end"))
    t = pa.Table.from_batches(out)
    assert t.num_rows == len(RANGES_TARGETS)  # exactly one row per probe row, in probe order
    rows = list(zip(*[t.column(i).to_pylist() for i in range(6)]))
    for row, probe in zip(rows, RANGES_TARGETS):
        assert row[3:] == probe
        hits = [a for a in RANGES_READS if a[0] == probe[0] and a[1] <= probe[2] and a[2] >= probe[1]]
        if hits:  # which overlapping interval is unpinned in the reference; here: smallest (start, end, row)
            assert row[:3] == min(hits, key=lambda a: (a[1], a[2]))
        else:
            assert row[:3] == (None, None, None)


# ---- TEST:608-657 count_overlaps, TEST:664-717 coverage (ranges fixtures; also with shared / separate column names) --------
@pytest.mark.parametrize("pos_type", [pa.int64(), pa.int32(), pa.uint32(), pa.uint64()])
@pytest.mark.parametrize("contig_type", [pa.string(), pa.large_string(), pa.string_view()])
def test_count_overlaps_and_coverage_udtf(ctx, pos_type, contig_type):
    init_ranges_tables(ctx, pos_type=pos_type, contig_type=contig_type)
    for args in ((), ("contig", "pos_start", "pos_end"), ("contig", "pos_start", "pos_end", "contig", "pos_start", "pos_end")):
        out = ctx.count_overlaps("reads", "targets", *args)
        t = pa.Table.from_batches(out)
        assert t.schema.names == ["contig", "pos_start", "pos_end", "count"]
        assert t.schema.field("count").type == pa.int64() and not t.schema.field("count").nullable
        assert t.column("count").to_pylist() == [2, 2, 2, 1, 1, 2, 2, 2, 1, 1, 0]
        assert t.column("pos_start").to_pylist() == [r[1] for r in RANGES_TARGETS]  # probe rows, probe order, untouched
        cov = pa.Table.from_batches(ctx.coverage("reads", "targets", *args))
        assert cov.schema.names[-1] == "coverage"
        assert cov.column("coverage").to_pylist() == [41, 92, 202, 1, 2, 41, 92, 202, 1, 2, 0]


# ---- TEST:1204-1252 strict filter op; one output batch per probe batch -----------------------------------------------------
def test_count_overlaps_strict_and_batching(ctx):
    ctx.register_table("l", _tbl([("chr1", 190, 300)]))
    ctx.register_table("r", _tbl([("chr1", 100, 190)] * 5), max_chunksize=2)
    assert pa.Table.from_batches(ctx.count_overlaps("l", "r")).column("count").to_pylist() == [1] * 5
    out = ctx.count_overlaps("l", "r", "contig", "pos_start", "pos_end", "strict")
    assert [b.num_rows for b in out] == [2, 2, 1]  # interval_tree.rs:236-273: one output batch per probe batch
    assert pa.Table.from_batches(out).column("count").to_pylist() == [0] * 5
    assert pa.Table.from_batches(ctx.count_overlaps("l", "r", "STRICT")).column("count").to_pylist() == [0] * 5


# ---- TEST:724-817 test_coverage_parquet: coverage(fBrain, exons, strict) == expected_coverage.parquet ------------------------
def test_coverage_golden_parquet_through_udtf(ctx):
    g = golden()
    contigs = [str(c) for c in g["contigs"]]

    def table(prefix):
        return pa.table({"contig": pa.array([contigs[i] for i in g[prefix + "_contig"]], pa.string()),
                         "pos_start": pa.array(g[prefix + "_start"], pa.int32()), "pos_end": pa.array(g[prefix + "_end"], pa.int32())})

    ctx.register_table("fbrain", table("fbrain"))
    ctx.register_table("exons", table("exons"), max_chunksize=8192)  # DataFusion's default batch size
    ctx.set_coalesce_rows(50_000)
    out = ctx.coverage("fbrain", "exons", "contig", "pos_start", "pos_end", "strict")
    assert len(out) == (438694 + 8191) // 8192
    cov = pa.Table.from_batches(out).column("coverage").to_numpy()
    assert cov.shape == (438694,) and int(cov.sum()) == 12060428
    assert (cov[g["exons_sort_order"]] == g["expected_coverage_sorted"]).all()


# ---- TEST:819-865 test_nearest_csv: CoitreesNearest join over the ranges fixtures ---------------------------------------------
def test_nearest_csv_join(ctx):
    init_ranges_tables(ctx)
    ctx.set("bio.interval_join_algorithm", "CoitreesNearest")
    out = ctx.interval_join("targets", "reads", [("contig", "contig")], ("pos_start", "pos_end"), ("pos_start", "pos_end"))
    t = pa.Table.from_batches(out)
    got = sorted(zip(*[t.column(i).to_pylist() for i in range(6)]))
    want = [
        ("chr1", 100, 190, "chr1", 150, 250), ("chr1", 100, 190, "chr1", 190, 300), ("chr1", 10000, 20000, "chr1", 15000, 15000),
        ("chr1", 22100, 22100, "chr1", 22000, 22300), ("chr1", 400, 600, "chr1", 300, 501), ("chr1", 400, 600, "chr1", 500, 700),
        ("chr2", 100, 190, "chr2", 150, 250), ("chr2", 100, 190, "chr2", 190, 300), ("chr2", 10000, 20000, "chr2", 15000, 15000),
        ("chr2", 22100, 22100, "chr2", 22000, 22300), ("chr2", 400, 600, "chr2", 300, 500), ("chr2", 400, 600, "chr2", 500, 700),
        ("chr3", 100, 200, "chr3", 234, 300),
    ]
    assert got == sorted(want)


# ---- TEST:1058-1103 test_bioframe_nearest_k1_schema_rows: nearest('reads','targets', 1, true) -----------------------------------
def test_nearest_udtf_k1_schema_and_rows(ctx):
    init_ranges_tables(ctx)
    out = ctx.nearest("reads", "targets", k=1, overlap=True)
    t = pa.Table.from_batches(out)
    assert t.schema.names == ["left_contig", "left_pos_start", "left_pos_end", "right_contig", "right_pos_start", "right_pos_end", "distance"]
    assert all(t.schema.field(n).nullable for n in t.schema.names)
    got = list(zip(*[t.column(n).to_pylist() for n in ["right_contig", "right_pos_start", "right_pos_end", "left_contig", "left_pos_start", "left_pos_end", "distance"]]))
    want = [
        ("chr1", 100, 190, "chr1", 150, 250, 0), ("chr1", 200, 290, "chr1", 150, 250, 0), ("chr1", 400, 600, "chr1", 300, 501, 0),
        ("chr1", 10000, 20000, "chr1", 15000, 15000, 0), ("chr1", 22100, 22100, "chr1", 22000, 22300, 0),
        ("chr2", 100, 190, "chr2", 150, 250, 0), ("chr2", 200, 290, "chr2", 150, 250, 0), ("chr2", 400, 600, "chr2", 300, 500, 0),
        ("chr2", 10000, 20000, "chr2", 15000, 15000, 0), ("chr2", 22100, 22100, "chr2", 22000, 22300, 0),
        ("chr3", 100, 200, "chr3", 234, 300, 34),
    ]
    assert got == want
    no_dist = pa.Table.from_batches(ctx.nearest("reads", "targets", compute_distance=False))
    assert "distance" not in no_dist.schema.names


# ---- TEST:905-949 nearest('l','r', 2, false): k > 1, overlaps excluded, NULL row for an unknown contig ------------------------------
def test_nearest_udtf_k2_no_overlap_and_null(ctx):
    ctx.register_table("l", _tbl([("chr1", 10, 20), ("chr1", 30, 40), ("chr1", 50, 60)]))
    ctx.register_table("r", _tbl([("chr1", 22, 22), ("chr1", 37, 37), ("chrX", 1, 1)]))
    t = pa.Table.from_batches(ctx.nearest("l", "r", k=2, overlap=False))
    got = sorted(zip(t.column("left_pos_start").to_pylist(), t.column("left_pos_end").to_pylist(),
                     t.column("right_pos_start").to_pylist(), t.column("distance").to_pylist()), key=repr)
    want = [(10, 20, 22, 2), (30, 40, 22, 8), (10, 20, 37, 17), (50, 60, 37, 13), (None, None, 1, None)]
    assert got == sorted(want, key=repr)


# ---- TEST:1254-1299 strict nearest: distance from raw coordinates; TEST:1301-1346 empty left table --------------------------------
def test_nearest_udtf_strict_distance_and_empty_left(ctx):
    ctx.register_table("l", _tbl([("chr1", 190, 190)]))
    ctx.register_table("r", _tbl([("chr1", 100, 190)]))
    t = pa.Table.from_batches(ctx.nearest("l", "r", "contig", "pos_start", "pos_end", "strict"))
    assert t.column("left_pos_start").to_pylist() == [190] and t.column("distance").to_pylist() == [0]
    ctx.register_table("e", _tbl([]))
    ctx.register_table("q", _tbl([("chr1", 100, 110), ("chr2", 200, 210)]))
    t = pa.Table.from_batches(ctx.nearest("e", "q"))
    assert t.num_rows == 2 and t.column("left_contig").to_pylist() == [None, None] and t.column("distance").to_pylist() == [None, None]
    assert t.column("right_pos_start").to_pylist() == [100, 200]


# ---- nearest.rs:323-324,346-352: the contig cache of the nearest loop starts at "" ---------------------------------------------------
def test_nearest_leading_empty_contig_quirk(ctx):
    ctx.register_table("l", _tbl([("", 10, 20), ("chr1", 10, 20)]))
    ctx.register_table("r", _tbl([("", 12, 13), ("chr1", 12, 13), ("", 12, 13)]))
    t = pa.Table.from_batches(ctx.nearest("l", "r"))
    # the leading "" row never consults the map; after "chr1" the same contig is looked up normally
    assert t.column("left_pos_start").to_pylist() == [None, 10, 10]
    c = pa.Table.from_batches(ctx.count_overlaps("l", "r"))  # count/coverage use Option<&str>: no such quirk
    assert c.column("count").to_pylist() == [1, 1, 1]


# ---- TEST:1612-1710, 1887-2061: overlap() UDTF, strictness, output modes ---------------------------------------------------------------
def test_overlap_udtf_modes(ctx):
    init_ranges_tables(ctx)
    t = pa.Table.from_batches(ctx.overlap("reads", "targets"))
    assert t.schema.names == ["left_contig", "left_pos_start", "left_pos_end", "right_contig", "right_pos_start", "right_pos_end"]
    want = sorted(a + b for a in RANGES_READS for b in RANGES_TARGETS if a[0] == b[0] and a[1] <= b[2] and a[2] >= b[1])
    assert sorted(zip(*[t.column(i).to_pylist() for i in range(6)])) == want
    ts = pa.Table.from_batches(ctx.overlap("reads", "targets", "strict"))
    want_s = sorted(a + b for a in RANGES_READS for b in RANGES_TARGETS if a[0] == b[0] and a[1] < b[2] and a[2] > b[1])
    assert sorted(zip(*[ts.column(i).to_pylist() for i in range(6)])) == want_s and len(want_s) < len(want)
    left_all = pa.Table.from_batches(ctx.overlap("reads", "targets", "left_all"))
    assert left_all.schema.names == ["contig", "pos_start", "pos_end"]
    assert sorted(zip(*[left_all.column(i).to_pylist() for i in range(3)])) == sorted(a for a in RANGES_READS for b in RANGES_TARGETS if a[0] == b[0] and a[1] <= b[2] and a[2] >= b[1])
    distinct = pa.Table.from_batches(ctx.overlap("reads", "targets", "left", "weak"))
    has = [a for a in RANGES_READS if any(a[0] == b[0] and a[1] <= b[2] and a[2] >= b[1] for b in RANGES_TARGETS)]
    assert list(zip(*[distinct.column(i).to_pylist() for i in range(3)])) == has  # RightSemi keeps probe order


# ---- RightSemi / RightAnti (interval_join.rs:1433-1463) and low-memory batching (interval_join.rs:1466-1573) ---------------------------------
def test_join_semi_anti_and_low_memory(ctx, monkeypatch):
    from bio_ranges_b200.session import JOIN_RIGHT_ANTI, JOIN_RIGHT_SEMI

    rng = np.random.default_rng(4)
    n = 4000
    def rnd(m):
        s = rng.integers(0, 50_000, m)
        return pa.table({"contig": pa.array([f"chr{int(c)}" for c in rng.integers(1, 4, m)]), "pos_start": pa.array(s, pa.int64()),
                         "pos_end": pa.array(s + rng.integers(0, 400, m), pa.int64())})
    a, b = rnd(n), rnd(3000)
    ctx.register_table("a", a)
    ctx.register_table("b", b, max_chunksize=1000)
    al, bl = list(zip(*[a.column(i).to_pylist() for i in range(3)])), list(zip(*[b.column(i).to_pylist() for i in range(3)]))
    import collections
    by = collections.defaultdict(list)
    for r in al:
        by[r[0]].append(r)
    match = [[x for x in by[y[0]] if x[1] <= y[2] and x[2] >= y[1]] for y in bl]
    on, cols = [("contig", "contig")], ("pos_start", "pos_end")
    semi = pa.Table.from_batches(ctx.interval_join("a", "b", on, cols, cols, join_type=JOIN_RIGHT_SEMI))
    anti = pa.Table.from_batches(ctx.interval_join("a", "b", on, cols, cols, join_type=JOIN_RIGHT_ANTI))
    assert semi.schema.names == ["contig", "pos_start", "pos_end"]
    assert list(zip(*[semi.column(i).to_pylist() for i in range(3)])) == [y for y, m in zip(bl, match) if m]
    assert list(zip(*[anti.column(i).to_pylist() for i in range(3)])) == [y for y, m in zip(bl, match) if not m]
    full = pa.Table.from_batches(ctx.interval_join("a", "b", on, cols, cols))
    want = sorted(x + y for y, m in zip(bl, match) for x in m)
    assert sorted(zip(*[full.column(i).to_pylist() for i in range(6)])) == want
    monkeypatch.setenv("BIO_MAX_OUTPUT_BATCH_SIZE", "500")
    ctx.set("bio.interval_join_low_memory", "true")
    out = ctx.interval_join("a", "b", on, cols, cols)
    assert len(out) > 3 and max(x.num_rows for x in out) <= max(500, max(len(m) for m in match))
    assert sorted(zip(*[pa.Table.from_batches(out).column(i).to_pylist() for i in range(6)])) == want


# ---- existence joins under the one-row-per-probe algorithms (interval_join.rs:849-870, 1433-1463): CoitreesCountOverlaps asks
#      the plain tree query, CoitreesNearest asks nearest_one(start, end, true) -- any probe row whose key has build rows matches
@pytest.mark.parametrize("alg", ["CoitreesCountOverlaps", "CoitreesNearest"])
def test_join_semi_anti_under_one_row_algorithms(ctx, alg):
    from bio_ranges_b200.session import JOIN_RIGHT_ANTI, JOIN_RIGHT_SEMI

    a = pa.table({"contig": ["chr1", "chr1", "chr2"], "pos_start": pa.array([10, 100, 5], pa.int64()), "pos_end": pa.array([20, 110, 9], pa.int64())})
    b = pa.table({"contig": ["chr1", "chr1", "chr2", "chr3", "chr1"], "pos_start": pa.array([15, 50, 100, 1, 200], pa.int64()),
                  "pos_end": pa.array([16, 60, 120, 2, 300], pa.int64())})
    ctx.register_table("a", a)
    ctx.register_table("b", b, max_chunksize=2)
    ctx.set("bio.interval_join_algorithm", alg)
    al, bl = list(zip(*[a.column(i).to_pylist() for i in range(3)])), list(zip(*[b.column(i).to_pylist() for i in range(3)]))
    if alg == "CoitreesNearest":  # nearest_one(.., true) is Some for every probe row of a key that has an index
        has = [any(x[0] == y[0] for x in al) for y in bl]
    else:
        has = [any(x[0] == y[0] and x[1] <= y[2] and x[2] >= y[1] for x in al) for y in bl]
    on, cols = [("contig", "contig")], ("pos_start", "pos_end")
    semi = _tab(ctx.interval_join("a", "b", on, cols, cols, join_type=JOIN_RIGHT_SEMI), b)
    anti = _tab(ctx.interval_join("a", "b", on, cols, cols, join_type=JOIN_RIGHT_ANTI), b)
    assert list(zip(*[semi.column(i).to_pylist() for i in range(3)])) == [y for y, m in zip(bl, has) if m]
    assert list(zip(*[anti.column(i).to_pylist() for i in range(3)])) == [y for y, m in zip(bl, has) if not m]
    # overlap(.., 'left') = RIGHT SEMI JOIN with the right table built, the left one probed (overlap.rs:205-226)
    ctx.register_table("reads", b)
    ctx.register_table("targets", a)
    d = _tab(ctx.overlap("reads", "targets", "left"), b)
    assert list(zip(*[d.column(i).to_pylist() for i in range(3)])) == [y for y, m in zip(bl, has) if m]


# ---- interval_join.rs:1963-2002: a coordinate that does not fit Int32 fails the join with arrow's cast error -------------------------------
def test_join_int32_cast_error(ctx):
    from bio_ranges_b200 import BioRangesError
    from bio_ranges_b200.session import ARROW, EXECUTION

    big = 2**31
    ctx.register_table("a", pa.table({"contig": ["a", "b"], "start": pa.array([5, 11], pa.int64()), "end": pa.array([10, 13], pa.int64())}))
    ctx.register_table("b", pa.table({"contig": ["a", "b"], "start": pa.array([1, 3], pa.int64()), "end": pa.array([2, big], pa.int64())}))
    with pytest.raises(BioRangesError) as ei:
        ctx.interval_join("a", "b", [("contig", "contig")], ("start", "end"), ("start", "end"))
    assert ei.value.status == ARROW and f"Arrow error: Cast error: Can't cast value {big} to type Int32" in str(ei.value)
    ctx.register_table("l", _tbl([("chr1", 1, 5)]))
    ctx.register_table("r", _tbl([("chr1", 1, big)]))
    with pytest.raises(BioRangesError) as ei:  # the table functions use PosArray::resolve's text (array_utils.rs:94-98)
        ctx.count_overlaps("l", "r")
    assert ei.value.status == EXECUTION and f"coordinate value {big} at row 0 overflows i32 (max 2147483647)" in str(ei.value)


# ---- Arrow ingest through the table functions: Int64 / UInt32 / UInt64 position columns are narrowed on the device -------------
def test_udtf_native_width_position_columns(ctx):
    from bio_ranges_b200 import BioRangesError
    from bio_ranges_b200.session import EXECUTION

    big = 2**31 + 5
    left = _tbl([("chr1", 100, 200), ("chr1", 150, 400), ("chr2", 10, 20)])
    ctx.register_table("l", left)
    want_i32 = None
    for ts, te in [(pa.int32(), pa.int32()), (pa.int64(), pa.int64()), (pa.uint32(), pa.uint64())]:
        r1 = pa.record_batch({"contig": ["chr1", "chr1"], "pos_start": pa.array([120, 500], ts), "pos_end": pa.array([160, 600], te)})
        r2 = pa.record_batch({"contig": ["chr2", "chr3"], "pos_start": pa.array([15, 1], ts), "pos_end": pa.array([30, 2], te)})
        ctx.register_table("r", pa.Table.from_batches([r1, r2]))
        got = (pa.Table.from_batches(ctx.count_overlaps("l", "r")).column("count").to_pylist(),
               pa.Table.from_batches(ctx.coverage("l", "r")).column("coverage").to_pylist(),
               pa.Table.from_batches(ctx.nearest("l", "r")).column("distance").to_pylist())
        if want_i32 is None:
            want_i32 = got
            assert got[0] == [2, 0, 1, 0]
        assert got == want_i32
    # the offending row is reported batch-local, in the first batch that holds one (the reference resolves batch by batch)
    r1 = pa.record_batch({"contig": ["chr1", "chr1"], "pos_start": pa.array([120, 500], pa.int64()), "pos_end": pa.array([160, 600], pa.int64())})
    r2 = pa.record_batch({"contig": ["chr2", "chr3"], "pos_start": pa.array([15, big], pa.int64()), "pos_end": pa.array([big + 1, 2], pa.int64())})
    ctx.register_table("r", pa.Table.from_batches([r1, r2]), max_chunksize=2)  # two probe batches
    for fn in (ctx.count_overlaps, ctx.coverage, ctx.nearest):
        with pytest.raises(BioRangesError) as ei:
            fn("l", "r")
        assert ei.value.status == EXECUTION and f"coordinate value {big} at row 1 overflows i32 (max 2147483647)" in str(ei.value)
