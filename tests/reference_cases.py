"""Known-answer cases transcribed from the reference's own tests, written once against a small backend
protocol so that BOTH the CPU oracle (tests/test_oracle_golden.py, -m "not gpu") and the CUDA path through the
C ABI (tests/test_gpu_parity.py, -m gpu) are pinned to the same literals.

Backend protocol: make_index(key, start, end, n_keys) -> object with
  count_overlaps(key,start,end,strict) -> int64[n]
  coverage(key,start,end,strict)       -> int64[n]
  nearest_k1(key,start,end,include_overlaps,strict,compute_distance) -> (left_idx u32[n], valid bool[n], dist i64[n]|None)
  nearest_k(key,start,end,k,include_overlaps,strict,compute_distance) -> (left_idx, valid, right_idx, dist)
  overlap_pairs(key,start,end) -> (build_idx u32[m], probe_idx u32[m])
Citations: IT = src/interval_tree.rs, NI = src/nearest_index.rs, TEST = tests/integration_test.rs,
IJ = src/physical_planner/joins/interval_join.rs (all under datafusion/bio-function-ranges/).
"""
import numpy as np

from _fixtures import (ABSENT, INTERVAL_READS, INTERVAL_TARGETS, RANGES_READS, RANGES_TARGETS, columns, encode, golden)


def _idx1(make_index, ivs):
    """single-key index from [(start,end[,pos])...]; build row id == list position"""
    s = np.array([iv[0] for iv in ivs], np.int32)
    e = np.array([iv[1] for iv in ivs], np.int32)
    return make_index(np.zeros(len(ivs), np.uint32), s, e, 1)


def _q(ix, fn, qs, qe, **kw):
    z = np.zeros(1, np.uint32)
    return getattr(ix, fn)(z, np.array([qs], np.int32), np.array([qe], np.int32), **kw)


# ---- IT:355-369 test_count_overlap_index_counts_inclusive_overlaps -------------------------------
def case_count_index_literals(make_index):
    ix = _idx1(make_index, [(10, 20), (15, 25), (30, 40)])
    for (qs, qe), want in [((5, 9), 0), ((5, 10), 1), ((20, 30), 3), ((26, 29), 0), ((31, 35), 1), ((10, 9), 0)]:
        assert _q(ix, "count_overlaps", qs, qe, strict=False)[0] == want, (qs, qe)


# ---- IT:371-412 get_coverage literals -------------------------------------------------------------
def case_coverage_literals(make_index):
    ix = _idx1(make_index, [(100, 200)])
    assert _q(ix, "coverage", 120, 150, strict=False)[0] == 32
    assert _q(ix, "coverage", 50, 120, strict=False)[0] == 21
    assert _q(ix, "coverage", 300, 400, strict=False)[0] == 0
    ix = _idx1(make_index, [(150, 250), (190, 300), (300, 501), (500, 700)])
    assert _q(ix, "coverage", 100, 190, strict=False)[0] == 41
    assert _q(ix, "coverage", 200, 290, strict=False)[0] == 92
    assert _q(ix, "coverage", 400, 600, strict=False)[0] == 202
    ix = _idx1(make_index, [(15000, 15000)])
    assert _q(ix, "coverage", 10000, 20000, strict=False)[0] == 1
    assert _q(ix, "coverage", 15000, 15000, strict=False)[0] == 1


# ---- NI:355-417 nearest unit tests (positions are explicit there; remap to row ids) ---------------
def _nearest_pos(make_index, recs, qs, qe, include_overlaps):
    ix = _idx1(make_index, recs)
    left, valid, _ = _q(ix, "nearest_k1", qs, qe, include_overlaps=include_overlaps, strict=False, compute_distance=True)
    return recs[int(left[0])][2] if valid[0] else None


def _nearest_k_pos(make_index, recs, qs, qe, k, include_overlaps):
    ix = _idx1(make_index, recs)
    left, valid, right, _ = _q(ix, "nearest_k", qs, qe, k=k, include_overlaps=include_overlaps, strict=False,
                               compute_distance=True)
    return [recs[int(l)][2] for l, v in zip(left, valid) if v]


def case_nearest_index_literals(make_index):
    assert _nearest_pos(make_index, [(0, 1000, 0), (900, 905, 1), (2000, 2010, 2)], 1100, 1100, True) == 0
    assert _nearest_pos(make_index, [(10, 20, 0), (30, 40, 1)], 12, 12, True) == 0
    # NI:371-377: positions 10,20,30 given in an order different from row order -> tie-break on (start,end,position)
    assert _nearest_pos(make_index, [(20, 30, 10), (10, 40, 20), (15, 25, 30)], 21, 21, True) == 20
    assert _nearest_pos(make_index, [(10, 20, 0), (10, 50, 1), (30, 40, 2)], 30, 30, True) == 1
    r4 = [(10, 20, 0), (30, 40, 1), (50, 60, 2), (70, 80, 3)]
    assert _nearest_k_pos(make_index, r4, 22, 22, 3, False) == [0, 1, 2]
    r3 = [(10, 20, 0), (30, 40, 1), (50, 60, 2)]
    assert _nearest_k_pos(make_index, r3, 35, 35, 2, True) == [1, 0]
    assert _nearest_k_pos(make_index, r3, 35, 35, 2, False) == [0, 2]
    # NI:409-416 empty index
    ix = make_index(np.zeros(0, np.uint32), np.zeros(0, np.int32), np.zeros(0, np.int32), 1)
    left, valid, right, _ = _q(ix, "nearest_k", 10, 20, k=3, include_overlaps=True, strict=False, compute_distance=True)
    assert list(valid) == [False]  # NE:433-439: one NULL row per probe row


# ---- TEST:608-657 / 1120-1143 / 1392-1428 count_overlaps(reads, targets) --------------------------
def case_count_overlaps_csv(make_index):
    bc, bs, be = columns(RANGES_READS)
    pc, ps, pe = columns(RANGES_TARGETS)
    bk, pk, nk = encode(bc, pc)
    ix = make_index(bk, bs, be, nk)
    got = ix.count_overlaps(pk, ps, pe, strict=False)
    assert list(got) == [2, 2, 2, 1, 1, 2, 2, 2, 1, 1, 0]


# ---- TEST:664-717 / 1159-1202 / 1439-1480 coverage(reads, targets) --------------------------------
def case_coverage_csv(make_index):
    bc, bs, be = columns(RANGES_READS)
    pc, ps, pe = columns(RANGES_TARGETS)
    bk, pk, nk = encode(bc, pc)
    ix = make_index(bk, bs, be, nk)
    got = ix.coverage(pk, ps, pe, strict=False)
    assert list(got) == [41, 92, 202, 1, 2, 41, 92, 202, 1, 2, 0]


# ---- TEST:1204-1252 strict boundary for count ------------------------------------------------------
def case_count_strict_boundary(make_index):
    ix = _idx1(make_index, [(190, 300)])
    assert _q(ix, "count_overlaps", 100, 190, strict=False)[0] == 1
    assert _q(ix, "count_overlaps", 100, 190, strict=True)[0] == 0
    # unknown contig -> 0 (IT:265)
    got = ix.count_overlaps(np.array([ABSENT], np.uint32), np.array([100], np.int32), np.array([400], np.int32), strict=False)
    assert got[0] == 0


# ---- TEST:724-817 test_coverage_parquet: coverage(fBrain -> exons, Strict) == golden --------------
def case_coverage_golden_parquet(make_index):
    g = golden()
    ix = make_index(g["fbrain_contig"].astype(np.uint32), g["fbrain_start"], g["fbrain_end"], len(g["contigs"]))
    cov = ix.coverage(g["exons_contig"].astype(np.uint32), g["exons_start"], g["exons_end"], strict=True)
    assert cov.shape == (438694,)
    got_sorted = cov[g["exons_sort_order"]]
    want = g["expected_coverage_sorted"].astype(np.int64)
    assert (got_sorted == want).all(), int((got_sorted != want).sum())
    assert int(cov.sum()) == 12060428


# ---- known answers derived from the fixtures: fBrain x exons pairs ---------------
def case_pairs_fbrain_exons(make_index):
    g = golden()
    ix = make_index(g["fbrain_contig"].astype(np.uint32), g["fbrain_start"], g["fbrain_end"], len(g["contigs"]))
    pk = g["exons_contig"].astype(np.uint32)
    b, p = ix.overlap_pairs(pk, g["exons_start"], g["exons_end"])
    assert len(b) == 54343
    assert len(np.unique(p)) == 51521
    assert np.bincount(p).max() == 7
    # strict join == `end - 1` on both sides (PI:63-124), i.e. the host passes adjusted columns
    ixs = make_index(g["fbrain_contig"].astype(np.uint32), g["fbrain_start"], g["fbrain_end"] - 1, len(g["contigs"]))
    b2, p2 = ixs.overlap_pairs(pk, g["exons_start"], g["exons_end"] - 1)
    assert len(b2) == 54246
    assert len(np.unique(p2)) == 51432
    # count_overlaps (weak) over the same data must sum to the weak pair count
    assert int(ix.count_overlaps(pk, g["exons_start"], g["exons_end"], strict=False).sum()) == 54343
    return (b, p), (b2, p2)


def _pairs_as_rows(build_rows, probe_rows, b, p):
    return sorted((build_rows[int(i)] + probe_rows[int(j)]) for i, j in zip(b, p))


# ---- TEST:58-137 test_equi_and_range_condition (build = reads = SQL-left, probe = targets) ---------
EXPECTED_EQUI = [
    ("chr1", 150, 250, "chr1", 100, 190), ("chr1", 150, 250, "chr1", 200, 290), ("chr1", 190, 300, "chr1", 100, 190),
    ("chr1", 190, 300, "chr1", 200, 290), ("chr1", 300, 501, "chr1", 400, 600), ("chr1", 500, 700, "chr1", 400, 600),
    ("chr1", 15000, 15000, "chr1", 10000, 20000), ("chr1", 22000, 22300, "chr1", 22100, 22100),
    ("chr2", 150, 250, "chr2", 100, 190), ("chr2", 150, 250, "chr2", 200, 290), ("chr2", 190, 300, "chr2", 100, 190),
    ("chr2", 190, 300, "chr2", 200, 290), ("chr2", 300, 500, "chr2", 400, 600), ("chr2", 500, 700, "chr2", 400, 600),
    ("chr2", 15000, 15000, "chr2", 10000, 20000), ("chr2", 22000, 22300, "chr2", 22100, 22100),
]


def case_join_equi(make_index):
    bc, bs, be = columns(INTERVAL_READS)
    pc, ps, pe = columns(INTERVAL_TARGETS)
    bk, pk, nk = encode(bc, pc)
    ix = make_index(bk, bs, be, nk)
    b, p = ix.overlap_pairs(pk, ps, pe)
    assert _pairs_as_rows(INTERVAL_READS, INTERVAL_TARGETS, b, p) == sorted(EXPECTED_EQUI)


# ---- TEST:139-231 test_range_condition: no equi key -> fake key on=[(1,1)] (PP:125-146): one group ---
def case_join_range_only(make_index):
    bc, bs, be = columns(INTERVAL_READS)
    pc, ps, pe = columns(INTERVAL_TARGETS)
    ix = make_index(np.zeros(len(bs), np.uint32), bs, be, 1)
    b, p = ix.overlap_pairs(np.zeros(len(ps), np.uint32), ps, pe)
    rows = _pairs_as_rows(INTERVAL_READS, INTERVAL_TARGETS, b, p)
    assert len(rows) == 32
    brute = sorted((INTERVAL_READS[i] + INTERVAL_TARGETS[j]) for i in range(12) for j in range(10)
                   if bs[i] <= pe[j] and be[i] >= ps[j])
    assert rows == brute


# ---- TEST:233-369 test_all_gteq_lteq_conditions / test_all_gt_lt_conditions ------------------------
_B12 = [(11, 15), (10, 15), (10, 10), (9, 15), (5, 15), (4, 15), (4, 10), (6, 8), (4, 8), (4, 5), (5, 5), (4, 4)]


def case_join_weak_vs_strict_tables(make_index):
    # q0: a is SQL-left (build, 1 row), b is probe
    ix = _idx1(make_index, [(5, 10)])
    ps = np.array([r[0] for r in _B12], np.int32)
    pe = np.array([r[1] for r in _B12], np.int32)
    z = np.zeros(len(_B12), np.uint32)
    b, p = ix.overlap_pairs(z, ps, pe)
    assert sorted(_B12[int(j)] for j in p) == sorted([(10, 15), (10, 10), (9, 15), (5, 15), (4, 15), (4, 10), (6, 8), (4, 8), (5, 5), (4, 5)])
    # strict: `<`/`>` => end - 1 on both sides (PI:85,105)
    ixs = _idx1(make_index, [(5, 10 - 1)])
    b, p = ixs.overlap_pairs(z, ps, pe - 1)
    assert sorted(_B12[int(j)] for j in p) == sorted([(9, 15), (5, 15), (4, 15), (4, 10), (6, 8), (4, 8)])
    # q1: b is SQL-left (build, 12 rows), a is probe -> same row set
    ixb = _idx1(make_index, [(s, e - 1) for s, e in _B12])
    b, p = ixb.overlap_pairs(np.zeros(1, np.uint32), np.array([5], np.int32), np.array([9], np.int32))
    assert sorted(_B12[int(i)] for i in b) == sorted([(9, 15), (5, 15), (4, 15), (4, 10), (6, 8), (4, 8)])


# ---- TEST:371-418 test_nearest (CoitreesNearest join, two-column key, strict) ----------------------
def case_join_coitrees_nearest(make_index):
    # build a = ('a','s',5,10); probe b rows; strict => end-1 on both sides; key = (contig,strand) tuple ids
    ix = make_index(np.zeros(1, np.uint32), np.array([5], np.int32), np.array([10 - 1], np.int32), 1)
    pk = np.array([0, 0, ABSENT, ABSENT], np.uint32)  # ('a','s'),('a','s'),('a','x'),('b','s')
    ps = np.array([11, 20, 0, 1], np.int32)
    pe = np.array([13, 21, 1, 2], np.int32) - 1
    left, valid, _ = ix.nearest_k1(pk, ps, pe, include_overlaps=True, strict=False, compute_distance=False)
    assert list(valid) == [True, True, False, False]
    assert list(left[:2]) == [0, 0]


# ---- TEST:819-865 test_nearest_csv (build = targets, probe = reads) ---------------------------------
def case_nearest_csv(make_index):
    bc, bs, be = columns(RANGES_TARGETS)
    pc, ps, pe = columns(RANGES_READS)
    bk, pk, nk = encode(bc, pc)
    ix = make_index(bk, bs, be, nk)
    left, valid, dist = ix.nearest_k1(pk, ps, pe, include_overlaps=True, strict=False, compute_distance=True)
    assert valid.all()
    got = sorted(RANGES_TARGETS[int(l)] + RANGES_READS[i] for i, l in enumerate(left))
    want = [
        ("chr1", 100, 190, "chr1", 150, 250), ("chr1", 100, 190, "chr1", 190, 300), ("chr1", 10000, 20000, "chr1", 15000, 15000),
        ("chr1", 22100, 22100, "chr1", 22000, 22300), ("chr1", 400, 600, "chr1", 300, 501), ("chr1", 400, 600, "chr1", 500, 700),
        ("chr2", 100, 190, "chr2", 150, 250), ("chr2", 100, 190, "chr2", 190, 300), ("chr2", 10000, 20000, "chr2", 15000, 15000),
        ("chr2", 22100, 22100, "chr2", 22000, 22300), ("chr2", 400, 600, "chr2", 300, 500), ("chr2", 400, 600, "chr2", 500, 700),
        ("chr3", 100, 200, "chr3", 234, 300),
    ]
    assert got == sorted(want)
    assert dist[-1] == 34 and (dist[:-1] == 0).all()


# ---- TEST:1058-1103 test_bioframe_nearest_k1_schema_rows: nearest('reads','targets',1,true) --------
def case_nearest_bioframe_k1(make_index):
    bc, bs, be = columns(RANGES_READS)
    pc, ps, pe = columns(RANGES_TARGETS)
    bk, pk, nk = encode(bc, pc)
    ix = make_index(bk, bs, be, nk)
    left, valid, dist = ix.nearest_k1(pk, ps, pe, include_overlaps=True, strict=False, compute_distance=True)
    assert valid.all()
    got = [RANGES_TARGETS[i] + RANGES_READS[int(l)] + (int(d),) for i, (l, d) in enumerate(zip(left, dist))]
    want = [
        ("chr1", 100, 190, "chr1", 150, 250, 0), ("chr1", 200, 290, "chr1", 150, 250, 0), ("chr1", 400, 600, "chr1", 300, 501, 0),
        ("chr1", 10000, 20000, "chr1", 15000, 15000, 0), ("chr1", 22100, 22100, "chr1", 22000, 22300, 0),
        ("chr2", 100, 190, "chr2", 150, 250, 0), ("chr2", 200, 290, "chr2", 150, 250, 0), ("chr2", 400, 600, "chr2", 300, 500, 0),
        ("chr2", 10000, 20000, "chr2", 15000, 15000, 0), ("chr2", 22100, 22100, "chr2", 22000, 22300, 0),
        ("chr3", 100, 200, "chr3", 234, 300, 34),
    ]
    assert got == want


# ---- TEST:905-949 nearest('l','r',2,false) incl. NULL row for unknown contig ------------------------
def case_nearest_k2_no_overlap_and_null(make_index):
    L = [(10, 20), (30, 40), (50, 60)]
    ix = _idx1(make_index, L)
    pk = np.array([0, 0, ABSENT], np.uint32)
    ps = np.array([22, 37, 1], np.int32)
    pe = np.array([22, 37, 1], np.int32)
    left, valid, right, dist = ix.nearest_k(pk, ps, pe, k=2, include_overlaps=False, strict=False, compute_distance=True)
    rows = [((L[int(l)] if v else None), int(r), (int(d) if v else None)) for l, v, r, d in zip(left, valid, right, dist)]
    want = [((10, 20), 0, 2), ((30, 40), 0, 8), ((10, 20), 1, 17), ((50, 60), 1, 13), (None, 2, None)]
    assert sorted(rows, key=repr) == sorted(want, key=repr)


# ---- TEST:1254-1299 strict nearest: search on adjusted coords, distance on RAW coords ---------------
def case_nearest_strict_distance_raw(make_index):
    ix = _idx1(make_index, [(190, 190)])
    left, valid, dist = _q(ix, "nearest_k1", 100, 190, include_overlaps=True, strict=True, compute_distance=True)
    assert valid[0] and left[0] == 0 and dist[0] == 0


# ---- TEST:1301-1346 empty build side -> one NULL row per probe row -----------------------------------
def case_nearest_empty_build(make_index):
    ix = make_index(np.zeros(0, np.uint32), np.zeros(0, np.int32), np.zeros(0, np.int32), 0)
    pk = np.array([ABSENT, ABSENT], np.uint32)
    left, valid, dist = ix.nearest_k1(pk, np.array([100, 200], np.int32), np.array([110, 210], np.int32),
                                      include_overlaps=True, strict=False, compute_distance=True)
    assert list(valid) == [False, False]
    c = ix.count_overlaps(pk, np.array([100, 200], np.int32), np.array([110, 210], np.int32), strict=False)
    assert list(c) == [0, 0]


# ---- TEST:1612-1710 overlap() strict / adjacency semantics ------------------------------------------
def case_overlap_adjacency(make_index):
    # overlap('reads','targets'): targets (R) is built, reads (L) probed (OV:173-180) -- symmetric predicate anyway
    def n_pairs(build, probe, strict):
        d = 1 if strict else 0
        ix = _idx1(make_index, [(build[0], build[1] - d)])
        b, p = ix.overlap_pairs(np.zeros(1, np.uint32), np.array([probe[0]], np.int32), np.array([probe[1] - d], np.int32))
        return len(b)

    assert n_pairs((100, 190), (190, 300), False) == 1
    assert n_pairs((100, 190), (190, 300), True) == 0
    assert n_pairs((200, 300), (100, 200), True) == 0
    assert n_pairs((200, 300), (100, 200), False) == 1
    assert n_pairs((100, 200), (100, 200), True) == 1


ALL_CASES = [v for k, v in sorted(globals().items()) if k.startswith("case_")]
