"""Parquet / Arrow ingest for the probe and build sides (replaces the host copies in front of the kernels:
PosArray::resolve, array_utils.rs:73-131, and the per-row contig hashing of build_*_from_batches, interval_tree.rs:75-143,
nearest.rs:498-547).

  * the contig column is read DICTIONARY-ENCODED straight from the Parquet dictionary pages (`read_dictionary`): no string
    is materialised per row; the few dictionary entries of a batch are mapped to build-side key ids once, and the per-row
    indices collapse to key RUNS (the reference's contig-cache loop, interval_tree.rs:258-265, on integers);
  * the position columns are handed to the C ABI as the Arrow value buffers they already are, in their native width
    (Int32 / Int64 / UInt32 / UInt64): the library moves them through its pinned staging ring and narrows + checks them on
    the device (brg_*_arrow), so the host never touches a coordinate.

Decoding Parquet is pyarrow's job here, as it is DataFusion's in the reference; everything after the decode is zero-copy on
the host side.
"""
import numpy as np

from . import _cabi as A


def _column_values(col, what):
    """the value buffer of a primitive column as a numpy view in its native width (nulls refused as the reference does,
    array_utils.rs:80-85)"""
    import pyarrow as pa

    if isinstance(col, pa.ChunkedArray):
        col = col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)
    if col.null_count:
        raise ValueError(f"{what} column contains null values")
    v = col.to_numpy(zero_copy_only=True)
    if v.dtype not in A.POS_TYPES:
        raise TypeError(f"unsupported {what} dtype {v.dtype}; expected Int32, Int64, UInt32 or UInt64")
    return v


def _dict_indices(col):
    import pyarrow as pa

    if isinstance(col, pa.ChunkedArray):
        col = col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)
    if not pa.types.is_dictionary(col.type):
        col = col.dictionary_encode()  # (a file whose writer did not dictionary-encode the column)
    if col.null_count:
        raise ValueError("contig column contains null values")
    return col.indices.to_numpy(zero_copy_only=False), col.dictionary.to_pylist()


def contig_runs(col, key_of):
    """dictionary-encoded contig column -> (run_key u32[r], run_end u64[r]); key_of: contig name -> build-side key id
    (names the build side does not have become KEY_ABSENT: count 0 / coverage 0 / NULL nearest)"""
    idx, names = _dict_indices(col)
    n = len(idx)
    if n == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint64)
    lut = np.array([key_of.get(s, A.KEY_ABSENT) for s in names], np.uint32)
    cut = np.flatnonzero(idx[1:] != idx[:-1]) + 1
    first = np.concatenate([[0], cut])
    return lut[idx[first]], np.concatenate([cut, [n]]).astype(np.uint64)


def iter_parquet(path, columns, batch_rows=1 << 22):
    """record batches of `columns` = (contig, start, end) with the contig column dictionary-encoded at scan time"""
    import pyarrow.parquet as pq

    pf = pq.ParquetFile(path, read_dictionary=[columns[0]])
    yield from pf.iter_batches(batch_size=batch_rows, columns=list(columns))


def build_from_parquet(ctx, path, columns=("contig", "pos_start", "pos_end"), want=A.WANT_ALL):
    """index over a Parquet interval table -> (Index, {contig name: key id}); key ids in first-seen order like the
    reference's grouping (interval_tree.rs:90-98)"""
    key_of = {}
    keys, starts, ends = [], [], []
    for rb in iter_parquet(path, columns):
        idx, names = _dict_indices(rb.column(0))
        lut = np.array([key_of.setdefault(s, len(key_of)) for s in names], np.uint32)
        keys.append(lut[idx])
        for dst, c, what in ((starts, rb.column(1), "start"), (ends, rb.column(2), "end")):
            v = _column_values(c, what)
            if v.dtype != np.int32:  # the build side is small: narrowed here, with the reference's check
                if len(v) and (v.max() > 2**31 - 1 or (v.dtype.kind == "i" and v.min() < -2**31)):
                    bad = int(np.flatnonzero((v > 2**31 - 1) | (v < -2**31))[0]) if v.dtype.kind == "i" else int(np.flatnonzero(v > 2**31 - 1)[0])
                    raise OverflowError(f"coordinate value {int(v[bad])} at row {bad} overflows i32 (max 2147483647)")
                v = v.astype(np.int32)
            dst.append(v)
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
    ix = ctx.build(cat(keys, np.uint32), cat(starts, np.int32), cat(ends, np.int32), max(1, len(key_of)), want)
    return ix, key_of


def probe_parquet(index, key_of, path, op, columns=("contig", "pos_start", "pos_end"), strict=False, include_overlaps=True,
                  batch_rows=1 << 22):
    """count_overlaps / coverage / nearest (k = 1) of every row of a Parquet table against `index`, batch by batch.
    op 'count' | 'coverage' -> int64[n];  'nearest' -> (left_idx u32[n], valid bool[n], distance i64[n])"""
    parts = []
    for rb in iter_parquet(path, columns, batch_rows):
        rk, rend = contig_runs(rb.column(0), key_of)
        if len(rend) == 0:
            continue
        s, e = _column_values(rb.column(1), "start"), _column_values(rb.column(2), "end")
        parts.append(index.probe_arrow(op, rk, rend, s, e, strict=strict, include_overlaps=include_overlaps))
    if op in ("count", "coverage"):
        return np.concatenate(parts) if parts else np.zeros(0, np.int64)
    if not parts:
        return np.zeros(0, np.uint32), np.zeros(0, bool), np.zeros(0, np.int64)
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))
