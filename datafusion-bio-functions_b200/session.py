"""BioSession: the reference's session-level API surface over the host layer of libbioranges.so (include/bioranges_host.h).

    ctx = create_bio_session()                                  # session_context.rs:140-148
    ctx.register_table("reads", pyarrow_table)                  # tests/integration_test.rs: ctx.register_csv / MemTable
    ctx.set("bio.interval_join_algorithm", "lapper")            # SET bio.interval_join_algorithm = lapper
    ctx.count_overlaps("reads", "targets", "contig", "pos_start", "pos_end", "strict")   # SELECT * FROM count_overlaps(...)
    ctx.coverage(...), ctx.nearest(...), ctx.overlap(...), ctx.interval_join(...)

Record batches cross the boundary through the Arrow C Data Interface (pyarrow's _export_to_c / _import_from_c, the same
structs arrow-rs exports as FFI_ArrowArray / FFI_ArrowSchema).  Everything below the call is C++ and CUDA; this file only
marshals arguments.
"""
import ctypes as C

from . import _cabi as A

PLAN, EXECUTION, NOT_IMPLEMENTED, INTERNAL, EXTERNAL, ARROW = 10, 11, 12, 13, 14, 15
JOIN_INNER, JOIN_RIGHT_SEMI, JOIN_RIGHT_ANTI = 0, 1, 2

HOST_SYMBOLS = [
    "brgh_create_bio_session", "brgh_session_destroy", "brgh_session_set", "brgh_session_get", "brgh_session_set_coalesce_rows",
    "brgh_register_batch", "brgh_deregister_table", "brgh_count_overlaps", "brgh_coverage", "brgh_nearest", "brgh_overlap",
    "brgh_merge", "brgh_cluster", "brgh_subtract", "brgh_complement", "brgh_interval_join", "brgh_parse_join_filter", "brgh_interval_join_filter", "brgh_explain_interval_join", "brgh_result_num_batches", "brgh_result_num_rows", "brgh_result_export",
    "brgh_result_release", "brgh_resolve_positions", "brgh_take",
]

class ColIntervals(C.Structure):
    """brgh_col_intervals (include/bioranges_host.h)"""
    _fields_ = [("left_start", C.c_char * 128), ("left_end", C.c_char * 128), ("right_start", C.c_char * 128),
                ("right_end", C.c_char * 128), ("left_end_minus_one", C.c_int), ("right_end_minus_one", C.c_int)]


_bound = False


def _lib():
    global _bound
    L = A.lib()
    if _bound:
        return L
    vp, u64, i32, i64, cp = C.c_void_p, C.c_uint64, C.c_int, C.c_int64, C.c_char_p
    strs = C.POINTER(cp)
    L.brgh_create_bio_session.argtypes = [i32, C.POINTER(vp)]
    L.brgh_session_destroy.argtypes = [vp]
    L.brgh_session_destroy.restype = None
    L.brgh_session_set.argtypes = [vp, cp, cp]
    L.brgh_session_get.argtypes = [vp, cp, cp, u64]
    L.brgh_session_set_coalesce_rows.argtypes = [vp, u64]
    L.brgh_register_batch.argtypes = [vp, cp, vp, vp]
    L.brgh_deregister_table.argtypes = [vp, cp]
    L.brgh_count_overlaps.argtypes = [vp, cp, cp, strs, i32, C.POINTER(vp)]
    L.brgh_coverage.argtypes = [vp, cp, cp, strs, i32, C.POINTER(vp)]
    L.brgh_nearest.argtypes = [vp, cp, cp, i64, i32, i32, strs, i32, C.POINTER(vp)]
    L.brgh_overlap.argtypes = [vp, cp, cp, strs, i32, C.POINTER(vp)]
    L.brgh_merge.argtypes = [vp, cp, i64, strs, i32, C.POINTER(vp)]
    L.brgh_cluster.argtypes = [vp, cp, i64, strs, i32, C.POINTER(vp)]
    L.brgh_subtract.argtypes = [vp, cp, cp, strs, i32, C.POINTER(vp)]
    L.brgh_complement.argtypes = [vp, cp, strs, i32, C.POINTER(vp)]
    L.brgh_interval_join.argtypes = [vp, cp, cp, strs, strs, i32, cp, cp, cp, cp, i32, i32, i32, C.POINTER(vp)]
    L.brgh_explain_interval_join.argtypes = [vp, i32, cp, u64]
    L.brgh_parse_join_filter.argtypes = [cp, cp, cp, C.POINTER(ColIntervals)]
    L.brgh_interval_join_filter.argtypes = [vp, cp, cp, strs, strs, i32, cp, cp, cp, i32, C.POINTER(vp)]
    L.brgh_result_num_batches.argtypes = [vp]
    L.brgh_result_num_batches.restype = u64
    L.brgh_result_num_rows.argtypes = [vp]
    L.brgh_result_num_rows.restype = u64
    L.brgh_result_export.argtypes = [vp, u64, vp, vp]
    L.brgh_result_release.argtypes = [vp]
    L.brgh_result_release.restype = None
    L.brgh_resolve_positions.argtypes = [vp, vp, cp, cp, vp, u64]
    L.brgh_take.argtypes = [vp, vp, vp, vp, u64, cp, vp, vp]
    _bound = True
    return L


def _cstrs(items):
    arr = (C.c_char_p * max(len(items), 1))()
    for i, s in enumerate(items):
        arr[i] = str(s).encode()
    return arr


def _export(batch):
    """pyarrow RecordBatch -> (ffi structs kept alive, array address, schema address)"""
    from pyarrow.cffi import ffi

    c_arr, c_sch = ffi.new("struct ArrowArray*"), ffi.new("struct ArrowSchema*")
    pa_, ps_ = int(ffi.cast("uintptr_t", c_arr)), int(ffi.cast("uintptr_t", c_sch))
    batch._export_to_c(pa_, ps_)
    return (c_arr, c_sch), pa_, ps_


def _import(call):
    """call(array_addr, schema_addr) fills the two structs; returns the pyarrow RecordBatch"""
    import pyarrow as pa
    from pyarrow.cffi import ffi

    c_arr, c_sch = ffi.new("struct ArrowArray*"), ffi.new("struct ArrowSchema*")
    pa_, ps_ = int(ffi.cast("uintptr_t", c_arr)), int(ffi.cast("uintptr_t", c_sch))
    call(pa_, ps_)
    return pa.RecordBatch._import_from_c(pa_, ps_)


class BioSession:
    def __init__(self, device=0):
        h = C.c_void_p()
        A.check(_lib().brgh_create_bio_session(int(device), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib().brgh_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- SET / SHOW bio.* ---------------------------------------------------------------------------------------------
    def set(self, key, value):
        A.check(_lib().brgh_session_set(self._h, key.encode(), str(value).encode()))

    def get(self, key):
        buf = C.create_string_buffer(256)
        A.check(_lib().brgh_session_get(self._h, key.encode(), buf, 256))
        return buf.value.decode()

    def set_coalesce_rows(self, rows):
        A.check(_lib().brgh_session_set_coalesce_rows(self._h, int(rows)))

    # ---- tables -------------------------------------------------------------------------------------------------------------
    def register_table(self, name, table, max_chunksize=None):
        """registers a pyarrow Table / RecordBatch as an in-memory table (one device probe batch per record batch)"""
        import pyarrow as pa

        _lib().brgh_deregister_table(self._h, name.encode())
        batches = [table] if isinstance(table, pa.RecordBatch) else table.combine_chunks().to_batches(max_chunksize=max_chunksize)
        if not batches:
            batches = [pa.RecordBatch.from_pylist([], schema=table.schema)]
        for b in batches:
            keep, pa_, ps_ = _export(b)
            A.check(_lib().brgh_register_batch(self._h, name.encode(), pa_, ps_))
            del keep

    # ---- table functions ------------------------------------------------------------------------------------------------------
    def _collect(self, res):
        import pyarrow as pa

        L = _lib()
        try:
            out = [_import(lambda a, s, i=i: A.check(L.brgh_result_export(res, i, a, s))) for i in range(L.brgh_result_num_batches(res))]
        finally:
            L.brgh_result_release(res)
        return out

    def _udtf(self, fn, left, right, args):
        res = C.c_void_p()
        A.check(fn(self._h, left.encode(), right.encode(), _cstrs(args), len(args), C.byref(res)))
        return self._collect(res)

    def count_overlaps(self, left, right, *args):
        """SELECT * FROM count_overlaps('left', 'right' [, cols...] [, 'strict'|'weak']) -> list of RecordBatch"""
        return self._udtf(_lib().brgh_count_overlaps, left, right, args)

    def coverage(self, left, right, *args):
        return self._udtf(_lib().brgh_coverage, left, right, args)

    def overlap(self, left, right, *args):
        return self._udtf(_lib().brgh_overlap, left, right, args)

    def subtract(self, left, right, *args):
        """SELECT * FROM subtract('left', 'right' [, cols...] [, 'strict'|'weak'])"""
        return self._udtf(_lib().brgh_subtract, left, right, args)

    def complement(self, table, *args):
        """SELECT * FROM complement('table' [, 'view_table'] [, col1, col2, col3 [, vcol1, vcol2, vcol3]] [, 'strict'|'weak'])"""
        res = C.c_void_p()
        A.check(_lib().brgh_complement(self._h, table.encode(), _cstrs(args), len(args), C.byref(res)))
        return self._collect(res)

    def merge(self, table, *args, min_dist=0):
        """SELECT * FROM merge('table' [, min_dist] [, col1, col2, col3] [, 'strict'|'weak'])"""
        res = C.c_void_p()
        A.check(_lib().brgh_merge(self._h, table.encode(), int(min_dist), _cstrs(args), len(args), C.byref(res)))
        return self._collect(res)

    def cluster(self, table, *args, min_dist=0):
        """SELECT * FROM cluster('table' [, min_dist] [, col1, col2, col3] [, 'strict'|'weak'])"""
        res = C.c_void_p()
        A.check(_lib().brgh_cluster(self._h, table.encode(), int(min_dist), _cstrs(args), len(args), C.byref(res)))
        return self._collect(res)

    def nearest(self, left, right, *args, k=1, overlap=True, compute_distance=True):
        res = C.c_void_p()
        A.check(_lib().brgh_nearest(self._h, left.encode(), right.encode(), int(k), int(bool(overlap)), int(bool(compute_distance)),
                                    _cstrs(args), len(args), C.byref(res)))
        return self._collect(res)

    def interval_join(self, left, right, on, left_cols, right_cols, lt_strict=False, gt_strict=False, join_type=JOIN_INNER):
        """SELECT * FROM left l JOIN right r ON l.on = r.on AND l.start <[=] r.end AND l.end >[=] r.start
        on: list of (left_col, right_col); left_cols / right_cols: (start, end)"""
        res = C.c_void_p()
        ol, orr = [a for a, _ in on], [b for _, b in on]
        A.check(_lib().brgh_interval_join(self._h, left.encode(), right.encode(), _cstrs(ol), _cstrs(orr), len(on),
                                          left_cols[0].encode(), left_cols[1].encode(), right_cols[0].encode(),
                                          right_cols[1].encode(), int(lt_strict), int(gt_strict), int(join_type), C.byref(res)))
        return self._collect(res)

    def interval_join_filter(self, left, right, on, filter, left_alias=None, right_alias=None, join_type=JOIN_INNER):
        """SELECT * FROM left <left_alias> JOIN right <right_alias> ON l.on = r.on AND <filter>: the range condition as text,
        parsed the way the planner rule parses the JoinFilter (physical_planner/intervals.rs)"""
        res = C.c_void_p()
        ol, orr = [a for a, _ in on], [b for _, b in on]
        A.check(_lib().brgh_interval_join_filter(self._h, left.encode(), right.encode(), _cstrs(ol), _cstrs(orr), len(on), filter.encode(),
                                                 (left_alias or left).encode(), (right_alias or right).encode(), int(join_type), C.byref(res)))
        return self._collect(res)

    def explain_interval_join(self, join_type=JOIN_INNER):
        buf = C.create_string_buffer(512)
        A.check(_lib().brgh_explain_interval_join(self._h, int(join_type), buf, 512))
        return buf.value.decode()


def parse_join_filter(filter, left_alias, right_alias):
    """the range-join filter parser alone (needs no GPU): -> dict(left_start, left_end, right_start, right_end,
    left_end_minus_one, right_end_minus_one)"""
    ci = ColIntervals()
    A.check(_lib().brgh_parse_join_filter(filter.encode(), left_alias.encode(), right_alias.encode(), C.byref(ci)))
    return {"left_start": ci.left_start.decode(), "left_end": ci.left_end.decode(), "right_start": ci.right_start.decode(),
            "right_end": ci.right_end.decode(), "left_end_minus_one": bool(ci.left_end_minus_one),
            "right_end_minus_one": bool(ci.right_end_minus_one)}


def create_bio_session(device=0):
    return BioSession(device)


def resolve_positions(batch, name, label="start"):
    """PosArray::resolve of one column of a pyarrow RecordBatch (array_utils.rs:73-131) -> numpy int32"""
    import numpy as np

    keep, pa_, ps_ = _export(batch)
    out = np.empty(batch.num_rows, np.int32)
    A.check(_lib().brgh_resolve_positions(pa_, ps_, name.encode(), label.encode(), out.ctypes.data, len(out)))
    return out


def take(batch, idx, valid=None, prefix=""):
    """arrow::compute::take of every column by a UInt32 index column with optional validity (bool array)"""
    import numpy as np

    idx = np.ascontiguousarray(idx, np.uint32)
    bits = None if valid is None else np.packbits(np.asarray(valid, bool), bitorder="little")
    keep, pa_, ps_ = _export(batch)
    return _import(lambda a, s: A.check(_lib().brgh_take(pa_, ps_, idx.ctypes.data, None if bits is None else bits.ctypes.data,
                                                         len(idx), prefix.encode(), a, s)))
