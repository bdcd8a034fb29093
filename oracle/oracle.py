"""ctypes loader for the CPU oracle (oracle/brg_oracle.c) and the reference-native cross-check
(oracle/_ref/libsi_ref.so, compiled from the reference's vendored superintervals.hpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ABSENT = 0xFFFFFFFF

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    """Compile liboracle.so (and _ref when /root/reference is present).  Building the checker is not using it."""
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "brg_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return so


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.brgo_build.restype = C.c_void_p
        L.brgo_build.argtypes = [C.c_void_p, _i32p, _i32p, C.c_uint64, C.c_uint32]
        L.brgo_index_destroy.argtypes = [C.c_void_p]
        for f in (L.brgo_count_overlaps, L.brgo_coverage):
            f.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_uint64, C.c_int, _i64p, C.c_int]
        L.brgo_nearest_k1.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_uint64, C.c_int, C.c_int,
                                      _u32p, _u8p, C.c_void_p, C.c_int]
        L.brgo_nearest_k.restype = C.c_uint64
        L.brgo_nearest_k.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int,
                                     _u32p, _u8p, _u32p, C.c_void_p]
        L.brgo_overlap_pairs.restype = C.c_uint64
        L.brgo_overlap_pairs.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_uint64,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.brgo_overlap_counts.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_uint64, _u64p, C.c_int]
        L.brgo_overlap_checksum.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_uint64, C.c_uint64, C.c_void_p, _u64p, C.c_int]
        L.brgo_overlap_pairs_partitions.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.brgo_overlap_pairs_partitions.restype = C.c_uint64
        L.brgo_merge_intervals.restype = C.c_uint64
        L.brgo_merge_intervals.argtypes = [_i32p, _i32p, C.c_uint64, _i32p, _i32p]
        L.brgo_free.argtypes = [C.c_void_p]
        L.brgo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _keyptr(key):
    if key is None:
        return None, None
    k = np.ascontiguousarray(key, dtype=np.uint32)
    return k, k.ctypes.data


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class OracleIndex:
    """Build-once / probe-many CPU index over (key_id, start, end) rows -- the reference's per-contig
    CountOverlapIndex / merged COITree / NearestIntervalIndex / join COITree all at once."""

    def __init__(self, key, start, end, n_keys=None):
        start, end = _i32(start), _i32(end)
        self._k, kp = _keyptr(key)
        if n_keys is None:
            n_keys = 1 if key is None else (int(self._k[self._k != ABSENT].max()) + 1 if len(self._k) and (self._k != ABSENT).any() else 0)
        self.n_keys = n_keys
        self.n = len(start)
        self._h = lib().brgo_build(kp, start, end, len(start), n_keys)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:  # (module globals are already gone at interpreter shutdown)
            try:
                _lib.brgo_index_destroy(self._h)
            except Exception:
                pass
            self._h = None

    def count_overlaps(self, key, start, end, strict=False, threads=1):
        start, end = _i32(start), _i32(end)
        k, kp = _keyptr(key)
        out = np.empty(len(start), np.int64)
        lib().brgo_count_overlaps(self._h, kp, start, end, len(start), int(strict), out, threads)
        return out

    def coverage(self, key, start, end, strict=False, threads=1):
        start, end = _i32(start), _i32(end)
        k, kp = _keyptr(key)
        out = np.empty(len(start), np.int64)
        lib().brgo_coverage(self._h, kp, start, end, len(start), int(strict), out, threads)
        return out

    def nearest_k1(self, key, start, end, include_overlaps=True, strict=False, compute_distance=True, threads=1):
        start, end = _i32(start), _i32(end)
        k, kp = _keyptr(key)
        n = len(start)
        left = np.empty(n, np.uint32)
        valid = np.empty(n, np.uint8)
        dist = np.empty(n, np.int64) if compute_distance else None
        lib().brgo_nearest_k1(self._h, kp, start, end, n, int(include_overlaps), int(strict), left, valid,
                              dist.ctypes.data if dist is not None else None, threads)
        return left, valid.astype(bool), dist

    def nearest_k(self, key, start, end, k, include_overlaps=True, strict=False, compute_distance=True):
        start, end = _i32(start), _i32(end)
        kk, kp = _keyptr(key)
        n = len(start)
        cap = max(1, n * max(k, 1))
        left = np.empty(cap, np.uint32)
        valid = np.empty(cap, np.uint8)
        right = np.empty(cap, np.uint32)
        dist = np.empty(cap, np.int64) if compute_distance else None
        w = lib().brgo_nearest_k(self._h, kp, start, end, n, k, int(include_overlaps), int(strict), left, valid, right,
                                 dist.ctypes.data if dist is not None else None)
        return left[:w].copy(), valid[:w].astype(bool), right[:w].copy(), (dist[:w].copy() if dist is not None else None)

    def overlap_pairs(self, key, start, end):
        """(build_idx, probe_idx) uint32 arrays, probe-major, match order within a probe unspecified."""
        start, end = _i32(start), _i32(end)
        k, kp = _keyptr(key)
        b, p = C.c_void_p(), C.c_void_p()
        m = lib().brgo_overlap_pairs(self._h, kp, start, end, len(start), C.byref(b), C.byref(p))
        bi = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint32)), shape=(m,)).copy() if m else np.empty(0, np.uint32)
        pi = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(m,)).copy() if m else np.empty(0, np.uint32)
        lib().brgo_free(b)
        lib().brgo_free(p)
        return bi, pi

    def overlap_counts(self, key, start, end, threads=1):
        start, end = _i32(start), _i32(end)
        k, kp = _keyptr(key)
        out = np.empty(len(start), np.uint64)
        lib().brgo_overlap_counts(self._h, kp, start, end, len(start), out, threads)
        return out


    def overlap_pairs_partitions(self, key, start, end, threads=1, row_base=0, checksum=False):
        """the pair join on `threads` partitions of the probe rows, every partition materialising its own (build_idx, probe_idx)
        vectors as the reference does; -> pair count, or (sum h, sum h^2, xor h, n_pairs) of the materialised vectors"""
        start, end = _i32(start), _i32(end)
        k, kp = _keyptr(key)
        sums = np.zeros(4, np.uint64) if checksum else None
        m = lib().brgo_overlap_pairs_partitions(self._h, kp, start, end, len(start), int(row_base), int(threads),
                                                sums.ctypes.data if checksum else None)
        return tuple(int(x) for x in sums) if checksum else int(m)

    def overlap_checksum(self, key, start, end, row_base=0, threads=1, counts=False):
        """order-independent checksum of the pair set: (sum h, sum h^2, xor h, n_pairs) as python ints, h = pair_hash(build_row,
        row_base + probe_row); with counts=True also the per-probe match counts"""
        start, end = _i32(start), _i32(end)
        k, kp = _keyptr(key)
        sums = np.zeros(4, np.uint64)
        cnt = np.empty(len(start), np.uint64) if counts else None
        lib().brgo_overlap_checksum(self._h, kp, start, end, len(start), int(row_base), cnt.ctypes.data if counts else None, sums, threads)
        t = tuple(int(x) for x in sums)
        return (t, cnt) if counts else t


def pair_hash(build_idx, probe_idx):
    """numpy twin of brgo_pair_hash (uint64 arithmetic wraps)"""
    b = np.asarray(build_idx).astype(np.uint64)
    p = np.asarray(probe_idx).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = b * np.uint64(0x9E3779B97F4A7C15) + p * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(0x165667B19E3779F9)
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(32)
    return h


def pair_checksum(build_idx, probe_idx):
    """(sum h, sum h^2, xor h, n) of a pair list, any order -- compare with OracleIndex.overlap_checksum"""
    h = pair_hash(build_idx, probe_idx)
    with np.errstate(over="ignore"):
        s1 = int(h.sum(dtype=np.uint64))
        s2 = int((h * h).sum(dtype=np.uint64))
    x = int(np.bitwise_xor.reduce(h)) if len(h) else 0
    return s1, s2, x, len(h)


def merge_intervals(first, last):
    first, last = _i32(first), _i32(last)
    of, ol = np.empty(max(1, len(first)), np.int32), np.empty(max(1, len(first)), np.int32)
    m = lib().brgo_merge_intervals(first, last, len(first), of, ol)
    return of[:m].copy(), ol[:m].copy()


def merge(key, start, end, n_keys, min_dist=0, strict=False):
    """merge UDTF core (merge.rs:260-358): returns (key u32[m], start i64[m], end i64[m], n_intervals i64[m])"""
    L = lib()
    L.brgo_merge.restype = C.c_uint64
    L.brgo_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_int, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_void_p]
    start = np.ascontiguousarray(start, np.int64)
    end = np.ascontiguousarray(end, np.int64)
    key = None if key is None else np.ascontiguousarray(key, np.uint32)
    n = len(start)
    ok, os_, oe, oc = np.empty(n, np.uint32), np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.int64)
    m = L.brgo_merge(None if key is None else key.ctypes.data, start.ctypes.data, end.ctypes.data, n, int(n_keys), int(min_dist),
                     int(bool(strict)), ok.ctypes.data, os_.ctypes.data, oe.ctypes.data, oc.ctypes.data)
    return ok[:m].copy(), os_[:m].copy(), oe[:m].copy(), oc[:m].copy()


def cluster(key, start, end, n_keys, min_dist=0, strict=False):
    """cluster UDTF core (cluster.rs:560-720): (row u64[n], cluster i64[n], cluster_start i64[n], cluster_end i64[n]) in
    (contig id, start, end, row) order"""
    L = lib()
    L.brgo_cluster.restype = None
    L.brgo_cluster.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p]
    start = np.ascontiguousarray(start, np.int64)
    end = np.ascontiguousarray(end, np.int64)
    key = None if key is None else np.ascontiguousarray(key, np.uint32)
    n = len(start)
    orow, oc, os_, oe = np.empty(n, np.uint64), np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.int64)
    L.brgo_cluster(None if key is None else key.ctypes.data, start.ctypes.data, end.ctypes.data, n, int(n_keys), int(min_dist),
                   int(bool(strict)), orow.ctypes.data, oc.ctypes.data, os_.ctypes.data, oe.ctypes.data)
    return orow, oc, os_, oe


def subtract(lkey, lstart, lend, rkey, rstart, rend, n_keys, strict=False):
    """subtract UDTF core (subtract.rs:395-440): (left_row u64[m], piece_start i64[m], piece_end i64[m]); left rows are walked
    in (key, start, end, row) order.  Keys are ids over the LEFT contigs; right rows with key >= n_keys are ignored."""
    L = lib()
    L.brgo_subtract.restype = C.c_uint64
    L.brgo_subtract.argtypes = [C.c_void_p] * 3 + [C.c_uint64] + [C.c_void_p] * 3 + [C.c_uint64, C.c_uint32, C.c_int] + [C.c_void_p] * 3 + [C.c_uint64]
    lkey, rkey = np.ascontiguousarray(lkey, np.uint32), np.ascontiguousarray(rkey, np.uint32)
    ls, le = np.ascontiguousarray(lstart, np.int64), np.ascontiguousarray(lend, np.int64)
    rs, re = np.ascontiguousarray(rstart, np.int64), np.ascontiguousarray(rend, np.int64)
    call = lambda o, a, b, cap: L.brgo_subtract(lkey.ctypes.data, ls.ctypes.data, le.ctypes.data, len(ls), rkey.ctypes.data, rs.ctypes.data,
                                                re.ctypes.data, len(rs), int(n_keys), int(bool(strict)), o, a, b, cap)
    m = call(None, None, None, 0)  # sizing pass
    orow, os_, oe = np.empty(m, np.uint64), np.empty(m, np.int64), np.empty(m, np.int64)
    call(orow.ctypes.data, os_.ctypes.data, oe.ctypes.data, m)
    return orow, os_, oe


def complement(key, start, end, n_keys, vkey, vstart, vend, strict=False):
    """complement UDTF core (complement.rs:297-358, 396-457): (view_row u64[m], gap_start i64[m], gap_end i64[m]), views
    answered in the order given; a view whose key has no input rows comes out whole."""
    L = lib()
    L.brgo_complement.restype = C.c_uint64
    L.brgo_complement.argtypes = [C.c_void_p] * 3 + [C.c_uint64, C.c_uint32] + [C.c_void_p] * 3 + [C.c_uint64, C.c_int] + [C.c_void_p] * 3 + [C.c_uint64]
    s, e = np.ascontiguousarray(start, np.int64), np.ascontiguousarray(end, np.int64)
    key = None if key is None else np.ascontiguousarray(key, np.uint32)
    vk = np.ascontiguousarray(vkey, np.uint32)
    vs, ve = np.ascontiguousarray(vstart, np.int64), np.ascontiguousarray(vend, np.int64)
    call = lambda o, a, b, cap: L.brgo_complement(None if key is None else key.ctypes.data, s.ctypes.data, e.ctypes.data, len(s), int(n_keys),
                                                  vk.ctypes.data, vs.ctypes.data, ve.ctypes.data, len(vk), int(bool(strict)), o, a, b, cap)
    m = call(None, None, None, 0)  # sizing pass
    ov, os_, oe = np.empty(m, np.uint64), np.empty(m, np.int64), np.empty(m, np.int64)
    call(ov.ctypes.data, os_.ctypes.data, oe.ctypes.data, m)
    return ov, os_, oe


def max_threads():
    """host cores this process may use.  NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers,
    which would silently turn the multi-threaded CPU baseline into a single-threaded one."""
    import os

    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


# ---------------------------------------------------------------------------------------------
# reference-native cross-check: the reference's vendored superintervals.hpp (oracle/_ref)
# ---------------------------------------------------------------------------------------------
def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libsi_ref.so"))


def ref():
    global _ref
    if _ref is None:
        R = C.CDLL(os.path.join(_HERE, "_ref", "libsi_ref.so"))
        R.si_ref_build.restype = C.c_void_p
        R.si_ref_build.argtypes = [C.c_void_p, _i32p, _i32p, C.c_uint64, C.c_uint32]
        R.si_ref_destroy.argtypes = [C.c_void_p]
        R.si_ref_counts.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_uint64, _u64p, C.c_int]
        R.si_ref_pairs.restype = C.c_uint64
        R.si_ref_pairs.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_uint64,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        R.si_ref_pairs_count_mt.restype = C.c_uint64
        R.si_ref_pairs_count_mt.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_uint64, C.c_int]
        R.si_ref_free.argtypes = [C.c_void_p]
        _ref = R
    return _ref


class SuperIntervalsRef:
    def __init__(self, key, start, end, n_keys):
        start, end = _i32(start), _i32(end)
        self._k, kp = _keyptr(key)
        self._h = ref().si_ref_build(kp, start, end, len(start), n_keys)

    def __del__(self):
        if getattr(self, "_h", None):
            ref().si_ref_destroy(self._h)
            self._h = None

    def counts(self, key, start, end, threads=1):
        start, end = _i32(start), _i32(end)
        k, kp = _keyptr(key)
        out = np.empty(len(start), np.uint64)
        ref().si_ref_counts(self._h, kp, start, end, len(start), out, threads)
        return out

    def pairs(self, key, start, end):
        start, end = _i32(start), _i32(end)
        k, kp = _keyptr(key)
        b, p = C.c_void_p(), C.c_void_p()
        m = ref().si_ref_pairs(self._h, kp, start, end, len(start), C.byref(b), C.byref(p))
        bi = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint32)), shape=(m,)).copy() if m else np.empty(0, np.uint32)
        pi = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(m,)).copy() if m else np.empty(0, np.uint32)
        ref().si_ref_free(b)
        ref().si_ref_free(p)
        return bi, pi

    def pairs_count_mt(self, key, start, end, threads):
        start, end = _i32(start), _i32(end)
        k, kp = _keyptr(key)
        return ref().si_ref_pairs_count_mt(self._h, kp, start, end, len(start), threads)


def canonical_pairs(build_idx, probe_idx):
    """Canonical sort for set comparison of pairs (match order within a probe is unspecified, SURVEY a'12)."""
    o = np.lexsort((build_idx, probe_idx))
    return np.asarray(build_idx)[o], np.asarray(probe_idx)[o]
