"""Object wrappers over the C ABI.  Host-buffer methods take numpy arrays (the caller-owned Arrow value buffers of
the reference's RecordBatches) and include the transfers; `*_dev` methods take torch CUDA tensors and enqueue on
torch's current stream (PyTorch is used for device memory and streams only)."""
import ctypes as C

import numpy as np

from . import _cabi as A


class Context:
    """One per GPU (one process per GPU): streams, staging buffers, scratch.  brg_ctx_create / brg_ctx_destroy."""

    def __init__(self, device=0):
        h = C.c_void_p()
        A.check(A.lib().brg_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    @property
    def sm_count(self):
        return A.lib().brg_ctx_sm_count(self._h)

    def close(self):
        if getattr(self, "_h", None):
            A.lib().brg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def take_dev(self, src, idx, out, valid_bits=None):
        """out[i] = src[idx[i]] on the device (fixed-width column gather by an emitted index column)"""
        import torch

        st = torch.cuda.current_stream().cuda_stream
        A.check(A.lib().brg_take_dev(self._h, _tptr(src), src.element_size(), _tptr(idx), _tptr(valid_bits), idx.numel(),
                                     _tptr(out), st))
        return out

    def sort_rows(self, key, start, end, n_keys):
        """row order by (key, start, end, row) -> uint32[n]"""
        key = A.as_key(key)
        start, end = A.as_i32(start), A.as_i32(end)
        out = np.empty(len(start), np.uint32)
        A.check(A.lib().brg_sort_rows(self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), len(start), int(n_keys), A.np_ptr(out)))
        return out

    def cluster(self, key, start, end, n_keys, min_dist=0, strict=False):
        """cluster UDTF core, one row per input row in (key, start, end, row) order:
        (row_id u32[n], cluster_start i32[n], cluster_end i32[n], cluster_id u32[n])"""
        return self.merge(key, start, end, n_keys, min_dist, strict, _fn="brg_cluster")

    def merge(self, key, start, end, n_keys, min_dist=0, strict=False, _fn="brg_merge"):
        """merge UDTF core: (key u32[m], start i32[m], end i32[m], n_intervals u32[m]) ordered by key id, then start"""
        key = A.as_key(key)
        start, end = A.as_i32(start), A.as_i32(end)
        h = C.c_void_p()
        A.check(getattr(A.lib(), _fn)(self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), len(start), int(n_keys), int(min_dist),
                                  int(bool(strict)), C.byref(h)))
        try:
            L = A.lib()
            m = L.brg_merged_len(h)
            out = (A.copy_from(L.brg_merged_key(h), m, np.uint32), A.copy_from(L.brg_merged_start(h), m, np.int32),
                   A.copy_from(L.brg_merged_end(h), m, np.int32), A.copy_from(L.brg_merged_count(h), m, np.uint32))
        finally:
            A.lib().brg_merged_release(h)
        return out

    def complement(self, key, start, end, n_keys, view_key, view_start, view_end, strict=False):
        """complement UDTF core: gaps of the merged input intervals inside each view [view_start, view_end) (i64), views
        answered in the order given -> (view_row u32[m], gap_start i64[m], gap_end i64[m])"""
        key = A.as_key(key)
        start, end = A.as_i32(start), A.as_i32(end)
        vk = np.ascontiguousarray(view_key, np.uint32)
        vs, ve = np.ascontiguousarray(view_start, np.int64), np.ascontiguousarray(view_end, np.int64)
        h = C.c_void_p()
        A.check(A.lib().brg_complement(self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), len(start), int(n_keys), A.np_ptr(vk),
                                       A.np_ptr(vs), A.np_ptr(ve), len(vk), int(bool(strict)), C.byref(h)))
        try:
            L = A.lib()
            m = L.brg_gaps_len(h)
            out = (A.copy_from(L.brg_gaps_view(h), m, np.uint32), A.copy_from(L.brg_gaps_start(h), m, np.int64),
                   A.copy_from(L.brg_gaps_end(h), m, np.int64))
        finally:
            A.lib().brg_gaps_release(h)
        return out

    # ---- build ------------------------------------------------------------------------------------------
    def build(self, key, start, end, n_keys, want=A.WANT_ALL):
        key = A.as_key(key)
        start, end = A.as_i32(start), A.as_i32(end)
        if len(start) != len(end) or (key is not None and len(key) != len(start)):
            raise ValueError("build columns differ in length")
        h = C.c_void_p()
        A.check(A.lib().brg_build(self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), len(start), int(n_keys),
                                  int(want), C.byref(h)))
        return Index(self, h)

    def build_dev(self, key, start, end, n_keys, want=A.WANT_ALL):
        import torch

        n = start.numel()
        h = C.c_void_p()
        st = torch.cuda.current_stream().cuda_stream
        A.check(A.lib().brg_build_dev(self._h, _tptr(key), _tptr(start), _tptr(end), n, int(n_keys), int(want), st,
                                      C.byref(h)))
        return Index(self, h)

    def import_index(self, header_bytes):
        """Allocates an empty replica slab for an index exported on another rank; returns (Index, slab_tensor)."""
        import torch

        h, slab, nbytes = C.c_void_p(), C.c_void_p(), C.c_uint64()
        buf = (C.c_char * len(header_bytes)).from_buffer_copy(header_bytes)
        A.check(A.lib().brg_index_import(self._h, buf, len(header_bytes), C.byref(h), C.byref(slab), C.byref(nbytes)))
        ix = Index(self, h)
        return ix, _wrap_device(slab.value, nbytes.value, self.device)


    def slab_alloc(self, nbytes):
        """an empty device buffer for a replica slab, as a torch uint8 tensor (receive the slab into it, then attach_slab)"""
        p = C.c_void_p()
        A.check(A.lib().brg_index_slab_alloc(self._h, int(nbytes), C.byref(p)))
        return p.value, _wrap_device(p.value, int(nbytes), self.device)

    def attach_slab(self, ptr, nbytes, stream=None):
        """Index over a received slab (it carries its own header); takes ownership of the buffer"""
        h = C.c_void_p()
        rc = A.lib().brg_index_attach(self._h, C.c_void_p(ptr), int(nbytes), stream, C.byref(h))
        if rc != 0:
            A.lib().brg_index_slab_free(self._h, C.c_void_p(ptr))
            A.check(rc)
        return Index(self, h)


def replicate(index, contexts):
    """replicas of `index` on the devices of `contexts`, all inside this process (brg_index_replicate)"""
    n = len(contexts)
    dst = (C.c_void_p * n)(*[c._h for c in contexts])
    out = (C.c_void_p * n)()
    A.check(A.lib().brg_index_replicate(index._h, dst, n, out))
    return [Index(c, C.c_void_p(h)) for c, h in zip(contexts, out)]


def _tptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


def _wrap_device(ptr, nbytes, device):
    """torch uint8 view of library-owned device memory (for torch.distributed.broadcast of the index slab)."""
    import torch

    class _Holder:
        pass

    holder = _Holder()
    holder.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(holder, device=f"cuda:{device}")


class Index:
    """Immutable build-side index resident in HBM (brg_index)."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self._h = handle

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            A.lib().brg_index_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def rows(self):
        return A.lib().brg_index_rows(self._h)

    @property
    def merged_rows(self):
        return A.lib().brg_index_merged_rows(self._h)

    @property
    def device_bytes(self):
        return A.lib().brg_index_device_bytes(self._h)

    def bins_info(self, table):
        """the bins table read for unsorted probe tables: table 0 = count, 1 = coverage, 2 = nearest / pairs.
        -> {"records", "bin_shift", "format"} (records == 0: the index holds no such table)"""
        sh, fmt = C.c_uint32(), C.c_uint32()
        n = A.lib().brg_index_bins_info(self._h, int(table), C.byref(sh), C.byref(fmt))
        return {"records": int(n), "bin_shift": int(sh.value), "format": int(fmt.value)}

    def export(self):
        """(header bytes, slab tensor) for replication to other GPUs."""
        n = C.c_uint64()
        slab, nbytes = C.c_void_p(), C.c_uint64()
        A.check(A.lib().brg_index_export(self._h, None, 0, C.byref(n), C.byref(slab), C.byref(nbytes)))
        buf = (C.c_char * n.value)()
        A.check(A.lib().brg_index_export(self._h, buf, n.value, C.byref(n), C.byref(slab), C.byref(nbytes)))
        return bytes(buf), _wrap_device(slab.value, nbytes.value, self.ctx.device)

    # ---- host-buffer probes (the drop-in calls) ---------------------------------------------------------------
    def _cols(self, key, start, end):
        key = A.as_key(key)
        start, end = A.as_i32(start), A.as_i32(end)
        if len(start) != len(end) or (key is not None and len(key) != len(start)):
            raise ValueError("probe columns differ in length")
        return key, start, end

    def count_overlaps(self, key, start, end, strict=False, out=None):
        key, start, end = self._cols(key, start, end)
        out = np.empty(len(start), np.int64) if out is None else out
        A.check(A.lib().brg_count_overlaps(self.ctx._h, self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), len(start),
                                           int(strict), A.np_ptr(out)))
        return out

    def count_overlaps_runs(self, run_key, run_end, start, end, strict=False, out=None):
        """count_overlaps with the probe key column in run-length form (see brg_count_overlaps_runs)"""
        start, end = A.as_i32(start), A.as_i32(end)
        run_key = np.ascontiguousarray(run_key, np.uint32)
        run_end = np.ascontiguousarray(run_end, np.uint64)
        out = np.empty(len(start), np.int64) if out is None else out
        A.check(A.lib().brg_count_overlaps_runs(self.ctx._h, self._h, A.np_ptr(run_key), A.np_ptr(run_end), len(run_key),
                                                A.np_ptr(start), A.np_ptr(end), len(start), int(strict), A.np_ptr(out)))
        return out

    def probe_arrow(self, op, run_key, run_end, start, end, strict=False, include_overlaps=True, compute_distance=True):
        """Arrow ingest (brg_*_arrow): start / end in their native width (int32 / int64 / uint32 / uint64), narrowed and
        checked on the device.  op: 'count' | 'coverage' | 'nearest'."""
        start, st = A.as_pos(start)
        end, et = A.as_pos(end)
        run_key = np.ascontiguousarray(run_key, np.uint32)
        run_end = np.ascontiguousarray(run_end, np.uint64)
        n = len(start)
        L = A.lib()
        if op in ("count", "coverage"):
            out = np.empty(n, np.int64)
            f = L.brg_count_overlaps_arrow if op == "count" else L.brg_coverage_arrow
            A.check(f(self.ctx._h, self._h, A.np_ptr(run_key), A.np_ptr(run_end), len(run_key), A.np_ptr(start), st, A.np_ptr(end), et, n,
                      int(strict), A.np_ptr(out), None))
            return out
        left = np.empty(n, np.uint32)
        bits = np.zeros(((n + 31) // 32) * 4, np.uint8)
        dist = np.empty(n, np.int64) if compute_distance else None
        A.check(L.brg_nearest_k1_arrow(self.ctx._h, self._h, A.np_ptr(run_key), A.np_ptr(run_end), len(run_key), A.np_ptr(start), st,
                                       A.np_ptr(end), et, n, int(include_overlaps), int(strict), A.np_ptr(left), A.np_ptr(bits),
                                       A.np_ptr(dist), None))
        return left, A.unpack_bits(bits, n), dist

    def coverage_runs(self, run_key, run_end, start, end, strict=False, out=None):
        start, end = A.as_i32(start), A.as_i32(end)
        run_key = np.ascontiguousarray(run_key, np.uint32)
        run_end = np.ascontiguousarray(run_end, np.uint64)
        out = np.empty(len(start), np.int64) if out is None else out
        A.check(A.lib().brg_coverage_runs(self.ctx._h, self._h, A.np_ptr(run_key), A.np_ptr(run_end), len(run_key),
                                          A.np_ptr(start), A.np_ptr(end), len(start), int(strict), A.np_ptr(out)))
        return out

    def nearest_k1_runs(self, run_key, run_end, start, end, include_overlaps=True, strict=False, compute_distance=True,
                        out=None):
        """out: optional (left_idx u32[n], valid_bits u8[4*ceil(n/32)], distance i64[n]) caller buffers (e.g. pinned);
        with `out` the packed Arrow validity bitmap is returned as is"""
        start, end = A.as_i32(start), A.as_i32(end)
        run_key = np.ascontiguousarray(run_key, np.uint32)
        run_end = np.ascontiguousarray(run_end, np.uint64)
        n = len(start)
        if out is None:
            left = np.empty(n, np.uint32)
            bits = np.zeros(((n + 31) // 32) * 4, np.uint8)
            dist = np.empty(n, np.int64) if compute_distance else None
        else:
            left, bits, dist = out
        A.check(A.lib().brg_nearest_k1_runs(self.ctx._h, self._h, A.np_ptr(run_key), A.np_ptr(run_end), len(run_key),
                                            A.np_ptr(start), A.np_ptr(end), n, int(include_overlaps), int(strict),
                                            A.np_ptr(left), A.np_ptr(bits), A.np_ptr(dist)))
        if out is not None:
            return left, bits, dist
        return left, A.unpack_bits(bits, n), dist

    def coverage(self, key, start, end, strict=False, out=None):
        key, start, end = self._cols(key, start, end)
        out = np.empty(len(start), np.int64) if out is None else out
        A.check(A.lib().brg_coverage(self.ctx._h, self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), len(start),
                                     int(strict), A.np_ptr(out)))
        return out

    def nearest_k1(self, key, start, end, include_overlaps=True, strict=False, compute_distance=True):
        key, start, end = self._cols(key, start, end)
        n = len(start)
        left = np.empty(n, np.uint32)
        bits = np.zeros(((n + 31) // 32) * 4, np.uint8)
        dist = np.empty(n, np.int64) if compute_distance else None
        A.check(A.lib().brg_nearest_k1(self.ctx._h, self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), n,
                                       int(include_overlaps), int(strict), A.np_ptr(left), A.np_ptr(bits), A.np_ptr(dist)))
        return left, A.unpack_bits(bits, n), dist

    def nearest_k1_into(self, key, start, end, out, include_overlaps=True, strict=False):
        """brg_nearest_k1 into caller buffers out = (left_idx u32[n], valid_bits u8[4*ceil(n/32)], distance i64[n] or None);
        the packed Arrow validity bitmap is returned as is"""
        key, start, end = self._cols(key, start, end)
        left, bits, dist = out
        A.check(A.lib().brg_nearest_k1(self.ctx._h, self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), len(start),
                                       int(include_overlaps), int(strict), A.np_ptr(left), A.np_ptr(bits), A.np_ptr(dist)))
        return out

    def nearest_k(self, key, start, end, k, include_overlaps=True, strict=False, compute_distance=True):
        key, start, end = self._cols(key, start, end)
        r = C.c_void_p()
        A.check(A.lib().brg_nearest(self.ctx._h, self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), len(start), int(k),
                                    int(include_overlaps), int(compute_distance), int(strict), C.byref(r)))
        try:
            L = A.lib()
            m = L.brg_rows_len(r)
            left = A.copy_from(L.brg_rows_build_idx(r), m, np.uint32)
            right = A.copy_from(L.brg_rows_probe_idx(r), m, np.uint32)
            bits = A.copy_from(L.brg_rows_valid_bits(r), ((m + 31) // 32) * 4, np.uint8)
            dist = A.copy_from(L.brg_rows_distance(r), m, np.int64) if compute_distance else None
        finally:
            A.lib().brg_rows_release(r)
        return left, A.unpack_bits(bits, m), right, dist

    def subtract(self, key, start, end, strict=False):
        """pieces of the given (left) intervals not covered by this (right, WANT_COVERAGE) index:
        (left_row u32[m], piece_start i32[m], piece_end i32[m]), left-row-major"""
        key, start, end = self._cols(key, start, end)
        r = C.c_void_p()
        A.check(A.lib().brg_subtract(self.ctx._h, self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), len(start), int(strict), C.byref(r)))
        try:
            L = A.lib()
            m = L.brg_rows_len(r)
            out = (A.copy_from(L.brg_rows_probe_idx(r), m, np.uint32), A.copy_from(L.brg_rows_piece_start(r), m, np.int32),
                   A.copy_from(L.brg_rows_piece_end(r), m, np.int32))
        finally:
            A.lib().brg_rows_release(r)
        return out

    def overlap_pairs(self, key, start, end, mode=A.PAIRS_INNER, max_rows=0, return_consumed=False, copy=True):
        """copy=False: the index arrays are views of the library's pinned result buffers (what a Rust host wraps with
        Buffer::from_custom_allocation); the buffers go back to the library when the last view is garbage-collected"""
        key, start, end = self._cols(key, start, end)
        r, consumed = C.c_void_p(), C.c_uint64()
        A.check(A.lib().brg_overlap_pairs(self.ctx._h, self._h, A.np_ptr(key), A.np_ptr(start), A.np_ptr(end), len(start),
                                          int(mode), int(max_rows), C.byref(r), C.byref(consumed)))
        if not copy:
            owner = _RowsOwner(r)
            L = A.lib()
            m = L.brg_rows_len(r)
            b = owner.view(L.brg_rows_build_idx(r), m, np.uint32) if mode in (A.PAIRS_INNER, A.PAIRS_INNER_ANY_ORDER) else None
            p = owner.view(L.brg_rows_probe_idx(r), m, np.uint32)
            if b is None:
                return (p, consumed.value) if return_consumed else p
            return (b, p, consumed.value) if return_consumed else (b, p)
        try:
            L = A.lib()
            m = L.brg_rows_len(r)
            b = A.copy_from(L.brg_rows_build_idx(r), m, np.uint32) if mode in (A.PAIRS_INNER, A.PAIRS_INNER_ANY_ORDER) else None
            p = A.copy_from(L.brg_rows_probe_idx(r), m, np.uint32)
        finally:
            A.lib().brg_rows_release(r)
        if b is None:
            return (p, consumed.value) if return_consumed else p
        return (b, p, consumed.value) if return_consumed else (b, p)

    # ---- device-buffer probes (inputs resident in HBM; enqueue on torch's current stream) ----------------------
    def count_overlaps_dev(self, key, start, end, out, strict=False):
        import torch

        st = torch.cuda.current_stream().cuda_stream
        A.check(A.lib().brg_count_overlaps_dev(self.ctx._h, self._h, _tptr(key), _tptr(start), _tptr(end), start.numel(),
                                               int(strict), _tptr(out), st))
        return out

    def coverage_dev(self, key, start, end, out, strict=False):
        import torch

        st = torch.cuda.current_stream().cuda_stream
        A.check(A.lib().brg_coverage_dev(self.ctx._h, self._h, _tptr(key), _tptr(start), _tptr(end), start.numel(),
                                         int(strict), _tptr(out), st))
        return out

    def nearest_k1_dev(self, key, start, end, left_idx, valid_bits, dist, include_overlaps=True, strict=False):
        import torch

        st = torch.cuda.current_stream().cuda_stream
        A.check(A.lib().brg_nearest_k1_dev(self.ctx._h, self._h, _tptr(key), _tptr(start), _tptr(end), start.numel(),
                                           int(include_overlaps), int(strict), _tptr(left_idx), _tptr(valid_bits),
                                           _tptr(dist), st))

    def pairs_count_dev(self, key, start, end, offsets):
        import torch

        st = torch.cuda.current_stream().cuda_stream
        A.check(A.lib().brg_pairs_count_dev(self.ctx._h, self._h, _tptr(key), _tptr(start), _tptr(end), start.numel(),
                                            _tptr(offsets), st))

    def pairs_emit_dev(self, key, start, end, offsets, build_idx, probe_idx):
        import torch

        st = torch.cuda.current_stream().cuda_stream
        A.check(A.lib().brg_pairs_emit_dev(self.ctx._h, self._h, _tptr(key), _tptr(start), _tptr(end), start.numel(),
                                           _tptr(offsets), _tptr(build_idx), _tptr(probe_idx), st))

    def pairs_chunk_offsets_dev(self, key, start, end, chunk_base, row_ranges):
        """pass 1.  chunk_base: int64[ceil(n/256)+1] ([c] = first output slot of 256-row chunk c, [-1] = total pairs);
        row_ranges: int64[n] (candidate range of every probe row, consumed by pairs_fill_dev)"""
        import torch

        st = torch.cuda.current_stream().cuda_stream
        A.check(A.lib().brg_pairs_chunk_offsets_dev(self.ctx._h, self._h, _tptr(key), _tptr(start), _tptr(end), start.numel(),
                                                    _tptr(chunk_base), _tptr(row_ranges), st))

    def pairs_fill_dev(self, start, chunk_base, row_ranges, build_idx, probe_idx):
        import torch

        st = torch.cuda.current_stream().cuda_stream
        A.check(A.lib().brg_pairs_fill_dev(self.ctx._h, self._h, _tptr(start), start.numel(), _tptr(chunk_base),
                                           _tptr(row_ranges), _tptr(build_idx), _tptr(probe_idx), st))

    def pairs_join_dev(self, key, start, end, build_idx, probe_idx, total, ordered=False, chunk_slots=None):
        """the join in one launch (brg_join.cu).  build_idx / probe_idx: int32[capacity] output buffers; total: int64[1] (device)
        receives the exact pair count -- pairs beyond the capacity are dropped.  ordered=False: the pairs of one chunk of
        brg_pairs_join_chunk_rows() probe rows are contiguous and probe-major, the chunks land in the order they finish;
        chunk_slots: optional int64[2 * n_chunks] = {first output slot, pairs} per chunk"""
        import torch

        st = torch.cuda.current_stream().cuda_stream
        cap = min(build_idx.numel(), probe_idx.numel())
        A.check(A.lib().brg_pairs_join_dev(self.ctx._h, self._h, _tptr(key), _tptr(start), _tptr(end), start.numel(), cap,
                                           A.JOIN_ORDERED if ordered else 0, _tptr(build_idx), _tptr(probe_idx), _tptr(total),
                                           _tptr(chunk_slots) if chunk_slots is not None else None, st))


class _RowsOwner:
    """keeps a brg_rows result alive while numpy views of its pinned buffers exist"""

    def __init__(self, handle):
        self._h = handle

    def view(self, ptr, n, dtype):
        if n == 0 or not ptr:
            return np.empty(0, dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        buf._brg_owner = self  # array.base -> buf -> owner
        return np.frombuffer(buf, dtype=dtype, count=n)

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                A.lib().brg_rows_release(self._h)
            except Exception:
                pass
            self._h = None


_default = {}


def default_context(device=0):
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
