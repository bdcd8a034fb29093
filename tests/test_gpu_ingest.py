"""Parquet ingest: the contig column arrives dictionary-encoded from the file, positions in their native
width; count / coverage / nearest equal the oracle run on the decoded columns."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _write(path, contigs, start, end, pos_type, row_group=50_000):
    t = pa.table({"contig": pa.array(contigs, pa.string()), "pos_start": pa.array(start, pos_type), "pos_end": pa.array(end, pos_type)})
    pq.write_table(t, path, row_group_size=row_group)


@pytest.mark.parametrize("pos_type", [pa.int64(), pa.int32(), pa.uint32()])
def test_parquet_ingest_equals_oracle(tmp_path, pos_type):
    import bio_ranges_b200 as brg
    from bio_ranges_b200 import ingest

    rng = np.random.default_rng(17)
    names = np.array(["chr1", "chr2", "chrX", "chrM", "chrUn_1"])
    nb, n = 40_000, 333_337
    bk = rng.integers(0, 4, nb)  # the build side has no chrUn_1
    bs = rng.integers(0, 5_000_000, nb)
    be = bs + rng.integers(0, 3000, nb)
    pk = np.sort(rng.integers(0, 5, n))
    ps = rng.integers(0, 5_000_000, n)
    o = np.lexsort((ps, pk))
    pk, ps = pk[o], ps[o]
    pe = ps + rng.integers(0, 400, n)
    _write(str(tmp_path / "targets.parquet"), names[bk], bs, be, pos_type)
    _write(str(tmp_path / "reads.parquet"), names[pk], ps, pe, pos_type)

    ctx = brg.Context(0)
    ix, key_of = ingest.build_from_parquet(ctx, str(tmp_path / "targets.parquet"))
    assert set(key_of) == set(names[:4]) and ix.rows == nb
    # oracle over the same key ids
    lut = np.array([key_of.get(s, O.ABSENT) for s in names], np.uint32)
    oi = O.OracleIndex(lut[bk], bs.astype(np.int32), be.astype(np.int32), len(key_of))
    qk, qs, qe = lut[pk], ps.astype(np.int32), pe.astype(np.int32)
    reads = str(tmp_path / "reads.parquet")
    for strict in (False, True):
        assert (ingest.probe_parquet(ix, key_of, reads, "count", strict=strict, batch_rows=100_000) == oi.count_overlaps(qk, qs, qe, strict=strict)).all()
        assert (ingest.probe_parquet(ix, key_of, reads, "coverage", strict=strict, batch_rows=70_001) == oi.coverage(qk, qs, qe, strict=strict)).all()
    gl, gv, gd = ingest.probe_parquet(ix, key_of, reads, "nearest")
    ol, ov, od = oi.nearest_k1(qk, qs, qe)
    assert (gv == ov).all() and (gl[ov] == ol[ov]).all() and (gd[ov] == od[ov]).all()
    assert not gv[pk == 4].any()  # a contig the build side does not have: NULL rows (nearest.rs:354-384)
    ix.close()
    ctx.close()


def test_parquet_ingest_reports_overflow(tmp_path):
    import bio_ranges_b200 as brg
    from bio_ranges_b200 import ingest

    _write(str(tmp_path / "t.parquet"), ["c"] * 3, [1, 5, 9], [2, 6, 10], pa.int64())
    _write(str(tmp_path / "r.parquet"), ["c"] * 4, [1, 2, 2**31 + 7, 4], [2, 3, 2**31 + 9, 5], pa.int64())
    ctx = brg.Context(0)
    ix, key_of = ingest.build_from_parquet(ctx, str(tmp_path / "t.parquet"), want=brg.WANT_COUNT)
    with pytest.raises(brg.BioRangesError, match=rf"coordinate value {2**31 + 7} at row 2 overflows i32 \(max 2147483647\)"):
        ingest.probe_parquet(ix, key_of, str(tmp_path / "r.parquet"), "count")
    ix.close()
    ctx.close()
