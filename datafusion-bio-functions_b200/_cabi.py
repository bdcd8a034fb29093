"""ctypes binding of include/bioranges.h (libbioranges.so).  This is the only way Python reaches the engine:
every call goes through the C ABI a Rust/DataFusion host would bind.  There is NO fallback: if the shared
library is missing or no sm_100 device is usable, the calls raise."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BRG_LIB") or os.path.join(_HERE, "libbioranges.so")  # BRG_LIB: developer switch (kernel variants)
KEY_ABSENT = 0xFFFFFFFF

WANT_COUNT, WANT_COVERAGE, WANT_NEAREST, WANT_PAIRS, WANT_ALL = 1, 2, 4, 8, 15
PAIRS_INNER, PAIRS_SEMI, PAIRS_ANTI, PAIRS_INNER_ANY_ORDER = 0, 1, 2, 3
JOIN_ORDERED = 1

# every symbol include/bioranges.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "brg_last_error", "brg_version", "brg_ctx_create", "brg_ctx_destroy", "brg_ctx_device", "brg_ctx_sm_count",
    "brg_pinned_alloc", "brg_pinned_free", "brg_build", "brg_build_dev", "brg_index_destroy", "brg_index_rows",
    "brg_index_merged_rows", "brg_index_device_bytes", "brg_index_bins_info", "brg_index_export", "brg_index_import", "brg_index_slab_alloc",
    "brg_index_slab_free", "brg_index_attach", "brg_index_replicate", "brg_count_overlaps",
    "brg_count_overlaps_dev", "brg_coverage", "brg_coverage_dev", "brg_nearest_k1", "brg_nearest_k1_dev", "brg_nearest",
    "brg_count_overlaps_runs", "brg_coverage_runs", "brg_nearest_k1_runs",
    "brg_count_overlaps_arrow", "brg_coverage_arrow", "brg_nearest_k1_arrow",
    "brg_overlap_pairs", "brg_pairs_count_dev", "brg_pairs_emit_dev", "brg_pairs_chunk_rows", "brg_pairs_chunk_offsets_dev",
    "brg_pairs_fill_dev", "brg_pairs_join_dev", "brg_pairs_join_chunk_rows", "brg_take_dev", "brg_merge", "brg_cluster", "brg_subtract", "brg_complement", "brg_gaps_len", "brg_gaps_view", "brg_gaps_start", "brg_gaps_end", "brg_gaps_release", "brg_rows_piece_start", "brg_rows_piece_end", "brg_sort_rows",
    "brg_merged_len", "brg_merged_key", "brg_merged_start",
    "brg_merged_end", "brg_merged_count", "brg_merged_release", "brg_rows_len", "brg_rows_build_idx",
    "brg_rows_probe_idx", "brg_rows_valid_bits", "brg_rows_distance", "brg_rows_release", "brg_kernel_launches",
]


POS_TYPES = {np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.uint32): 2, np.dtype(np.uint64): 3}  # BRG_POS_*


def as_pos(a):
    """a position column in its Arrow width (Int32 / Int64 / UInt32 / UInt64) -> (contiguous array, BRG_POS_* code)"""
    a = np.ascontiguousarray(a)
    if a.dtype not in POS_TYPES:
        raise TypeError(f"unsupported position dtype {a.dtype}; expected int32, int64, uint32 or uint64")
    return a, POS_TYPES[a.dtype]


class BioRangesError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libbioranges status {status}: {msg}")
        self.status = status


_lib = None


def lib():
    """Loads libbioranges.so (built in-tree by `make -C datafusion-bio-functions_b200` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension must be built (python -c 'import __graft_entry__ as g; g.build()')."
            " There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.brg_last_error.restype = C.c_char_p
    L.brg_version.restype = i32
    L.brg_kernel_launches.restype = u64
    L.brg_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.brg_ctx_destroy.argtypes = [vp]
    L.brg_ctx_destroy.restype = None
    L.brg_ctx_device.argtypes = [vp]
    L.brg_ctx_sm_count.argtypes = [vp]
    L.brg_pinned_alloc.argtypes = [u64, C.POINTER(vp)]
    L.brg_pinned_free.argtypes = [vp]
    L.brg_pinned_free.restype = None
    L.brg_build.argtypes = [vp, vp, vp, vp, u64, u32, u32, C.POINTER(vp)]
    L.brg_build_dev.argtypes = [vp, vp, vp, vp, u64, u32, u32, vp, C.POINTER(vp)]
    L.brg_index_destroy.argtypes = [vp]
    L.brg_index_destroy.restype = None
    for f in (L.brg_index_rows, L.brg_index_merged_rows, L.brg_index_device_bytes):
        f.argtypes = [vp]
        f.restype = u64
    L.brg_index_bins_info.argtypes = [vp, i32, C.POINTER(u32), C.POINTER(u32)]
    L.brg_index_bins_info.restype = u64
    L.brg_index_export.argtypes = [vp, vp, u64, C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)]
    L.brg_index_import.argtypes = [vp, vp, u64, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]
    L.brg_index_slab_alloc.argtypes = [vp, u64, C.POINTER(vp)]
    L.brg_index_slab_free.argtypes = [vp, vp]
    L.brg_index_slab_free.restype = None
    L.brg_index_attach.argtypes = [vp, vp, u64, vp, C.POINTER(vp)]
    L.brg_index_replicate.argtypes = [vp, C.POINTER(vp), i32, C.POINTER(vp)]
    L.brg_count_overlaps.argtypes = [vp, vp, vp, vp, vp, u64, i32, vp]
    L.brg_coverage.argtypes = [vp, vp, vp, vp, vp, u64, i32, vp]
    L.brg_count_overlaps_runs.argtypes = [vp, vp, vp, vp, u64, vp, vp, u64, i32, vp]
    L.brg_coverage_runs.argtypes = [vp, vp, vp, vp, u64, vp, vp, u64, i32, vp]
    L.brg_nearest_k1_runs.argtypes = [vp, vp, vp, vp, u64, vp, vp, u64, i32, i32, vp, vp, vp]
    L.brg_count_overlaps_arrow.argtypes = [vp, vp, vp, vp, u64, vp, i32, vp, i32, u64, i32, vp, vp]
    L.brg_coverage_arrow.argtypes = [vp, vp, vp, vp, u64, vp, i32, vp, i32, u64, i32, vp, vp]
    L.brg_nearest_k1_arrow.argtypes = [vp, vp, vp, vp, u64, vp, i32, vp, i32, u64, i32, i32, vp, vp, vp, vp]
    L.brg_count_overlaps_dev.argtypes = [vp, vp, vp, vp, vp, u64, i32, vp, vp]
    L.brg_coverage_dev.argtypes = [vp, vp, vp, vp, vp, u64, i32, vp, vp]
    L.brg_nearest_k1.argtypes = [vp, vp, vp, vp, vp, u64, i32, i32, vp, vp, vp]
    L.brg_nearest_k1_dev.argtypes = [vp, vp, vp, vp, vp, u64, i32, i32, vp, vp, vp, vp]
    L.brg_nearest.argtypes = [vp, vp, vp, vp, vp, u64, u32, i32, i32, i32, C.POINTER(vp)]
    L.brg_overlap_pairs.argtypes = [vp, vp, vp, vp, vp, u64, i32, u64, C.POINTER(vp), C.POINTER(u64)]
    L.brg_pairs_count_dev.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp]
    L.brg_pairs_emit_dev.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp, vp, vp]
    L.brg_pairs_chunk_rows.restype = u32
    L.brg_pairs_chunk_offsets_dev.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp, vp]
    L.brg_pairs_fill_dev.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp, vp]
    L.brg_pairs_join_dev.argtypes = [vp, vp, vp, vp, vp, u64, u64, i32, vp, vp, vp, vp, vp]
    L.brg_pairs_join_chunk_rows.restype = u32
    L.brg_take_dev.argtypes = [vp, vp, u32, vp, vp, u64, vp, vp]
    L.brg_merge.argtypes = [vp, vp, vp, vp, u64, u32, C.c_int64, i32, C.POINTER(vp)]
    L.brg_cluster.argtypes = [vp, vp, vp, vp, u64, u32, C.c_int64, i32, C.POINTER(vp)]
    L.brg_subtract.argtypes = [vp, vp, vp, vp, vp, u64, i32, C.POINTER(vp)]
    L.brg_sort_rows.argtypes = [vp, vp, vp, vp, u64, u32, vp]
    L.brg_complement.argtypes = [vp, vp, vp, vp, u64, u32, vp, vp, vp, u64, i32, C.POINTER(vp)]
    L.brg_gaps_len.argtypes = [vp]
    L.brg_gaps_len.restype = u64
    for f in (L.brg_gaps_view, L.brg_gaps_start, L.brg_gaps_end):
        f.argtypes = [vp]
        f.restype = vp
    L.brg_gaps_release.argtypes = [vp]
    L.brg_gaps_release.restype = None
    for f in (L.brg_rows_piece_start, L.brg_rows_piece_end):
        f.argtypes = [vp]
        f.restype = vp
    L.brg_merged_len.argtypes = [vp]
    L.brg_merged_len.restype = u64
    for f in (L.brg_merged_key, L.brg_merged_start, L.brg_merged_end, L.brg_merged_count):
        f.argtypes = [vp]
        f.restype = vp
    L.brg_merged_release.argtypes = [vp]
    L.brg_merged_release.restype = None
    L.brg_rows_len.argtypes = [vp]
    L.brg_rows_len.restype = u64
    for f in (L.brg_rows_build_idx, L.brg_rows_probe_idx, L.brg_rows_valid_bits, L.brg_rows_distance):
        f.argtypes = [vp]
        f.restype = vp
    L.brg_rows_release.argtypes = [vp]
    L.brg_rows_release.restype = None
    _lib = L
    return L


def check(status):
    if status != 0:
        raise BioRangesError(status, lib().brg_last_error().decode("utf-8", "replace"))


def np_ptr(a):
    return None if a is None else a.ctypes.data


def as_i32(a):
    a = np.asarray(a)
    if a.dtype != np.int32:
        raise TypeError(f"coordinate column must be int32 at the C ABI (got {a.dtype}); resolve it on the host first")
    return np.ascontiguousarray(a)


def as_key(a):
    if a is None:
        return None
    return np.ascontiguousarray(np.asarray(a), dtype=np.uint32)


def key_runs(key):
    """run-length form of a per-row key column: (run_key u32[r], run_end u64[r]) -- what the host's contig-cache loop
    produces while it walks a coordinate-sorted table"""
    key = np.ascontiguousarray(key, dtype=np.uint32)
    n = len(key)
    if n == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint64)
    cut = np.flatnonzero(key[1:] != key[:-1]) + 1
    starts = np.concatenate([[0], cut])
    return key[starts].copy(), np.concatenate([cut, [n]]).astype(np.uint64)


def copy_from(ptr, n, dtype):
    """copy n elements of dtype out of library-owned memory"""
    if n == 0 or not ptr:
        return np.empty(0, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


def unpack_bits(bits_u8, n):
    return np.unpackbits(bits_u8, bitorder="little")[:n].astype(bool)
