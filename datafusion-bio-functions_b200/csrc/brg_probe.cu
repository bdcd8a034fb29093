// brg_probe.cu -- K3: probe kernels (count_overlaps, coverage, nearest) for sm_100a.
//
// Every probe row is independent given the read-only index.  The kernels stream the probe columns once
// (touch-once loads, evict-first), answer each row with rank queries through the rank directories (one 32 B
// KeyDir sector + two adjacent directory words + a search confined to one small bin of the sorted array, all
// L2/L1 resident), and stream the result columns once.  Nothing is a contraction: no tensor cores.
//
//   count_overlaps : #(starts <= qe) - #(ends < qs)                       interval_tree.rs:41-49
//   coverage       : clipped boundary terms + prefix-sum of interior merged widths
//                                                                         interval_tree.rs:145-152
//   nearest (k=1)  : first overlap by (start,end,pos) via prefix_max_end, else nearest left/right neighbour
//                                                                         nearest_index.rs:91-101,192-266
//   nearest (k>1)  : overlaps in (start,end,pos) order, then a two-pointer walk outwards
//                                                                         nearest_index.rs:103-190
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "brg_internal.cuh"
#include "brg_tile.cuh"
#include "brg_win.cuh"

namespace brg {

constexpr int PB_THREADS = 256;
#ifndef BRG_NEAREST_PREFETCH
#define BRG_NEAREST_PREFETCH 1
#endif
#ifndef BRG_NEAREST_THREADS
#define BRG_NEAREST_THREADS 256
#endif
#ifndef BRG_COV_COMPACT
#define BRG_COV_COMPACT 1
#endif
constexpr int COV_LIST_CAP = 192;  // rows of a chunk the compacted coverage epilogue lists (3 x 192 words = aux1 + aux2)
#ifndef BRG_COVERAGE_THREADS
#define BRG_COVERAGE_THREADS 320  // as BRG_COUNT_THREADS: 96 registers, no spills; 0.521 -> 0.491 ms per 100M probes
#endif
#ifndef BRG_COUNT_THREADS
#define BRG_COUNT_THREADS 320  // threads per CTA of the tiled count kernel: 96 registers, no spills, 20 warps per SM (0.549 -> 0.532 ms;
                               // 352 / 384 threads spill: 0.61 / 0.58 ms).  env BRG_COUNT_THREADS overrides (debugging switch)
#endif
#ifndef BRG_CC_CTAS
#define BRG_CC_CTAS 2  // resident CTAs per SM of the count / coverage tiled kernels: 128 registers per thread, no spills
                       // (3 CTAs = 80 registers spilled in the hot loop: 0.73 ms vs 0.59 ms per 125M probes on B200)
#endif

__device__ __forceinline__ uint32_t load_key(const uint32_t* key, uint64_t i) { return key ? ld_stream(key + i) : 0u; }

// ---------------------------------------------------------------------------------------------------
// count_overlaps
// ---------------------------------------------------------------------------------------------------
// #(starts <= qe) - #(ends < qs) of one row of key c: through the compact index when the build made one.
// skeys = the compact index' per-key constants in shared memory ([0] starts, [n_keys] ends), or nullptr.
constexpr uint32_t CI_SMEM_KEYS = 256;  // keys whose constants the per-row kernels keep in shared memory (2 x 8 KB)
__device__ __forceinline__ int64_t count_one_row(const IndexView& ix, const CompactKey* skeys, uint32_t c, int32_t qs, int32_t qe) {
    if (ix.cb.rec) {  // count bins: one 32-byte record answers both rank queries when the row lies inside one bin (the usual case)
        const KeyDir k = skeys ? reinterpret_cast<const KeyDir*>(skeys)[c] : load_keydir(ix.cb.kd, c);
        if (k.nb == 0) return 0;
        uint32_t started, ended;
        const uint32_t missing = bins_pair_ranks(ix.cb, k, qe, qs, &started, &ended);
        if (missing & 1u) started = rank_le_kd(ix.d_sstart, load_keydir(ix.d_sstart.kd, c), qe);
        if (missing & 2u) ended = rank_lt_kd(ix.d_eend, load_keydir(ix.d_eend.kd, c), qs);
        return (int64_t)started - (int64_t)ended;
    }
    if (ix.c_sstart.blocks) {
        if (skeys)
            return (int64_t)rank_le_compact(ix.c_sstart, skeys[c], qe) - (int64_t)rank_lt_compact(ix.c_eend, skeys[ix.n_keys + c], qs);
        return (int64_t)rank_le_compact(ix.c_sstart, load_compact_key(ix.c_sstart, c), qe) -
               (int64_t)rank_lt_compact(ix.c_eend, load_compact_key(ix.c_eend, c), qs);
    }
    const KeyDir ks = load_keydir(ix.d_sstart.kd, c);
    const KeyDir ke = load_keydir(ix.d_eend.kd, c);
    return (int64_t)rank_le_kd(ix.d_sstart, ks, qe) - (int64_t)rank_lt_kd(ix.d_eend, ke, qs);  // (both ranks include seg_begin)
}
// copies the compact index' per-key constants to shared memory (all threads of the CTA; returns nullptr when they do not fit)
__device__ __forceinline__ const CompactKey* stage_compact_keys(const IndexView& ix, CompactKey* smem) {
    if (ix.cb.rec) {  // count bins: one 32-byte KeyDir per key
        if (ix.n_keys > 2 * CI_SMEM_KEYS) return nullptr;
        const uint32_t words = ix.n_keys * (uint32_t)(sizeof(KeyDir) / 4);
        const uint32_t* g = reinterpret_cast<const uint32_t*>(ix.cb.kd);
        uint32_t* s = reinterpret_cast<uint32_t*>(smem);
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) s[i] = __ldg(g + i);
        __syncthreads();
        return smem;
    }
    if (!ix.c_sstart.blocks || ix.n_keys > CI_SMEM_KEYS) return nullptr;
    const uint32_t words = ix.n_keys * (uint32_t)(sizeof(CompactKey) / 4);
    const uint32_t* g0 = reinterpret_cast<const uint32_t*>(ix.c_sstart.d.kd);
    const uint32_t* g1 = reinterpret_cast<const uint32_t*>(ix.c_eend.d.kd);
    uint32_t* s = reinterpret_cast<uint32_t*>(smem);
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) { s[i] = __ldg(g0 + i); s[words + i] = __ldg(g1 + i); }
    __syncthreads();
    return smem;
}

__global__ void __launch_bounds__(PB_THREADS) count_kernel(IndexView ix, const uint32_t* __restrict__ key,
                                                           const int32_t* __restrict__ start,
                                                           const int32_t* __restrict__ end, uint64_t n, int strict,
                                                           int64_t* __restrict__ out) {
    __shared__ CompactKey skeys_buf[2 * CI_SMEM_KEYS];
    const CompactKey* skeys = stage_compact_keys(ix, skeys_buf);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t qs = ld_stream(start + i), qe = ld_stream(end + i);
        uint32_t c = load_key(key, i);
        if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); }  // interval_tree.rs:253-256
        int64_t cnt = 0;
        if (c < ix.n_keys && !(qe < qs)) {                    // unknown contig -> 0 (:265); guard (:42-44)
            cnt = count_one_row(ix, skeys, c, qs, qe);
        }
        st_stream(out + i, cnt);
    }
}

// Tiled variant (the hot path): TMA-staged probe tiles + warp-window rank search, see brg_tile.cuh.
__device__ __forceinline__ void store8_i64(int64_t* out, const long long (&v)[PV]) {
    longlong2* o = reinterpret_cast<longlong2*>(out);
#pragma unroll
    for (int j = 0; j < PV / 2; j++) __stcs(o + j, make_longlong2(v[2 * j], v[2 * j + 1]));
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, BRG_CC_CTAS) count_kernel_tiled(IndexView ix, const uint32_t* __restrict__ key,
                                                                 const int32_t* __restrict__ start,
                                                                 const int32_t* __restrict__ end, uint64_t n_rows,
                                                                 int strict, int64_t* __restrict__ out, TodoList todo) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpRing* rings = reinterpret_cast<WarpRing*>(smem_raw);  // one per warp of the CTA
    uint32_t* aux0 = rings[threadIdx.x >> 5].aux[0];
    uint32_t* aux1 = rings[threadIdx.x >> 5].aux[1];
    ProbePipe pipe(rings, key, start, end, n_rows);
    ProbeTile t;
    WarpCursor cs, ce;
    cs.reset();
    ce.reset();
    KeyDir ks, ke;
    uint32_t kd_key = 0xFFFFFFFFu;
    uint32_t fails = 0;  // consecutive deferred chunks
    while (pipe.next(t)) {
        if (strict) strict_adjust(t);
        pad_tile(t);
        uint32_t kuni;
        bool ok = warp_uniform_key(t, ix.n_keys, &kuni);
        uint32_t started[PV], ended[PV];
        if (ok) {
            if (kuni != kd_key) {
                kd_key = kuni;
                ks = load_keydir(ix.d_sstart.kd, kuni);
                ke = load_keydir(ix.d_eend.kd, kuni);
            }
            // seg_begin + #(starts <= qe)  and  seg_begin + #(ends < qs)
            const bool mono_s = chunk_mono(t.qs, t.full), mono_e = chunk_mono(t.qe, t.full);
            // both searches interleaved when both windows are short (the common case), else one after the other
            if (!(mono_s && mono_e &&
                  warp_window_ranks2<false, true>(cs, ix.d_sstart, ks, t.qe, t.s_qe, aux0, started, ce, ix.d_eend, ke, t.qs, t.s_qs,
                                                  aux1, ended, kuni))) {
                ok = warp_ranks_le_end(cs, ix.d_sstart, ks, kuni, t.qs, t.qe, t.active, mono_s, mono_e, t.s_qs, t.s_qe, aux0, started);
                ok = ok && warp_window_ranks<true>(ce, ix.d_eend, ke, kuni, t.qs, t.active, mono_s, t.s_qs, aux1, ended);
            }
        }
        if (ok) {
            // end < start -> 0 (interval_tree.rs:42-44); otherwise (started - ended_before) as i64 (:46-48), which is negative
            // when build rows with start > end make more ends than starts precede the probe.  With fewer than 2^31 build rows
            // the difference is exact as a signed 32-bit value, which is what the transposed store widens.
            uint32_t c32[PV];
#pragma unroll
            for (int j = 0; j < PV; j++) c32[j] = (t.qe[j] < t.qs[j]) ? 0u : started[j] - ended[j];
            if (t.full && ix.n < 0x80000000u) {
                store_chunk_i64_from32(aux0, out + (t.row0 - (threadIdx.x & 31) * PV), c32, true);
            } else if (t.active) {
                long long cnt[PV];
#pragma unroll
                for (int j = 0; j < PV; j++) cnt[j] = (t.qe[j] < t.qs[j]) ? 0ll : (long long)started[j] - (long long)ended[j];
                store8_i64(out + t.row0, cnt);
            }
            fails = 0;
        } else {
            // mixed keys (contig boundary, unsorted input), keys absent from the build side, or a window wider than the
            // shared buffer: hand the chunk to the per-row kernel
            cs.valid = false;
            ce.valid = false;
            todo_push(todo, t.chunk);
            fails++;
        }
        if (fails >= TP_BAIL || cs.wide || ce.wide) { pipe.bail_out(todo); break; }
        pipe.release();
    }
}

// deferred chunks: one row per thread, rank directory, full occupancy (memory-level parallelism for random probes)
__global__ void __launch_bounds__(TP_WTILE) count_kernel_todo(IndexView ix, const uint32_t* __restrict__ key,
                                                              const int32_t* __restrict__ start,
                                                              const int32_t* __restrict__ end, uint64_t n_rows, int strict,
                                                              int64_t* __restrict__ out, TodoList todo) {
    const uint32_t n_todo = *todo.count;
    if (blockIdx.x >= n_todo) return;  // (nothing was deferred: skip staging the key constants)
    __shared__ CompactKey skeys_buf[2 * CI_SMEM_KEYS];
    const CompactKey* skeys = stage_compact_keys(ix, skeys_buf);
    for (uint32_t c = blockIdx.x; c < n_todo; c += gridDim.x) {
        const uint64_t i = (uint64_t)todo.chunks[c] * TP_WTILE + threadIdx.x;
        if (i >= n_rows) continue;
        int32_t qs = ld_stream(start + i), qe = ld_stream(end + i);
        const uint32_t k = load_key(key, i);
        if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); }
        int64_t cnt = 0;
        if (k < ix.n_keys && !(qe < qs)) cnt = count_one_row(ix, skeys, k, qs, qe);
        st_stream(out + i, cnt);
    }
}

// ---------------------------------------------------------------------------------------------------
// coverage
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t cov_term(int32_t qs, int32_t qe, int32_t first, int32_t last) {
    int32_t hi = min(wadd(qe, 1), last);   // min(end + 1, node.last)
    int32_t lo = max(wsub(qs, 1), first);  // max(start - 1, node.first)
    int32_t ov = wsub(hi, lo);
    return ov < 1 ? 1 : ov;                // max(1, ...)
}

template <bool EARLY_OUT>
__device__ __forceinline__ int32_t coverage_from_range(const IndexView& ix, uint32_t lo, uint32_t hi, int32_t qs, int32_t qe);

// [lo, hi) = the merged intervals of key c that overlap [qs, qe]: lo = first j with last_j >= qs, hi = first j with first_j > qe.
// Through the coverage bins (one 32-byte record for both when the row lies inside a bin) when the build made them, else through
// the two rank directories.  skeys = the bins' per-key constants in shared memory, or nullptr.
__device__ __forceinline__ void coverage_row_range(const IndexView& ix, const KeyDir* skeys, uint32_t c, int32_t qs, int32_t qe,
                                                   uint32_t* lo, uint32_t* hi) {
    if (ix.cvb.rec) {
        const KeyDir k = skeys ? skeys[c] : load_keydir(ix.cvb.kd, c);
        if (k.nb == 0) { *lo = *hi = 0; return; }
        const uint32_t missing = bins_pair_ranks(ix.cvb, k, qe, qs, hi, lo);
        if (missing & 1u) *hi = rank_le_kd(ix.d_mfirst, load_keydir(ix.d_mfirst.kd, c), qe);
        if (missing & 2u) *lo = rank_lt_kd(ix.d_mlast, load_keydir(ix.d_mlast.kd, c), qs);
        return;
    }
    *lo = rank_lt_kd(ix.d_mlast, load_keydir(ix.d_mlast.kd, c), qs);
    *hi = rank_le_kd(ix.d_mfirst, load_keydir(ix.d_mfirst.kd, c), qe);
}
__device__ __forceinline__ const KeyDir* stage_coverage_keys(const IndexView& ix, KeyDir* smem) {
    if (!ix.cvb.rec || ix.n_keys > CI_SMEM_KEYS) return nullptr;
    const uint32_t words = ix.n_keys * (uint32_t)(sizeof(KeyDir) / 4);
    const uint32_t* g = reinterpret_cast<const uint32_t*>(ix.cvb.kd);
    uint32_t* s = reinterpret_cast<uint32_t*>(smem);
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) s[i] = __ldg(g + i);
    __syncthreads();
    return smem;
}

__global__ void __launch_bounds__(PB_THREADS) coverage_kernel(IndexView ix, const uint32_t* __restrict__ key,
                                                              const int32_t* __restrict__ start,
                                                              const int32_t* __restrict__ end, uint64_t n, int strict,
                                                              int64_t* __restrict__ out) {
    __shared__ KeyDir skeys_buf[CI_SMEM_KEYS];
    const KeyDir* skeys = stage_coverage_keys(ix, skeys_buf);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t qs = ld_stream(start + i), qe = ld_stream(end + i);
        uint32_t c = load_key(key, i);
        if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); }
        int32_t acc = 0;
        if (c < ix.n_keys) {
            // merged intervals are disjoint and ascending in both first and last: the ones overlapping [qs,qe]
            // (first <= qe && last >= qs, the coitrees predicate) are exactly the index range [lo, hi)
            uint32_t lo, hi;
            coverage_row_range(ix, skeys, c, qs, qe, &lo, &hi);
            acc = coverage_from_range<true>(ix, lo, hi, qs, qe);
        }
        st_stream(out + i, (int64_t)acc);  // interval_tree.rs:207
    }
}

// one row given the index range [lo, hi) of merged intervals that overlap it (interval_tree.rs:145-152).
// The two boundary intervals get the clipped formula, the interior ones lie strictly inside (qs-1, qe+1) so their term
// is max(1, last-first) = a difference of the width prefix.  m_rec[j] = {first, last, pre[j], pre[j+1]}: two 128-bit
// loads, no data-dependent branch on the common path.
template <bool EARLY_OUT>
__device__ __forceinline__ int32_t coverage_from_range(const IndexView& ix, uint32_t lo, uint32_t hi, int32_t qs, int32_t qe) {
    if (EARLY_OUT && lo >= hi) return 0;  // per-row kernels: most reads of a sparse target set touch nothing
    if (qs == INT32_MIN || qe == INT32_MAX) {
        // `start - 1` / `end + 1` wrap in the reference's release arithmetic: keep the literal per-interval sum
        int32_t acc = 0;
        for (uint32_t j = lo; j < hi; j++) {
            const int4 m = __ldg(ix.m_rec + j);
            acc = wadd(acc, cov_term(qs, qe, m.x, m.y));
        }
        return acc;
    }
    // predicated loads: a read that touches no target (most reads of a sparse target set) issues neither, a read inside one
    // merged interval only the first -- fewer wavefronts through the load/store pipe, no branch
    int4 a = make_int4(0, 0, 0, 0);
    if (lo < hi) a = __ldg(ix.m_rec + lo);
    int4 b = a;
    if (hi > lo + 1) b = __ldg(ix.m_rec + hi - 1);
    int32_t acc = cov_term(qs, qe, a.x, a.y);
    const int32_t rest = wadd(cov_term(qs, qe, b.x, b.y), wsub(b.z, a.w));
    acc = (hi > lo + 1) ? wadd(acc, rest) : acc;
    return (lo < hi) ? acc : 0;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, BRG_CC_CTAS) coverage_kernel_tiled(IndexView ix, const uint32_t* __restrict__ key,
                                                                    const int32_t* __restrict__ start,
                                                                    const int32_t* __restrict__ end, uint64_t n_rows,
                                                                    int strict, int64_t* __restrict__ out, TodoList todo) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpRing* rings = reinterpret_cast<WarpRing*>(smem_raw);  // one per warp of the CTA
    uint32_t* aux0 = rings[threadIdx.x >> 5].aux[0];
    uint32_t* aux1 = rings[threadIdx.x >> 5].aux[1];
    ProbePipe pipe(rings, key, start, end, n_rows);
    ProbeTile t;
    WarpCursor cl, cf;
    cl.reset();
    cf.reset();
    KeyDir kl, kf;
    uint32_t kd_key = 0xFFFFFFFFu;
    uint32_t fails = 0;  // consecutive deferred chunks
    while (pipe.next(t)) {
        if (strict) strict_adjust(t);
        pad_tile(t);
        uint32_t kuni;
        bool ok = warp_uniform_key(t, ix.n_keys, &kuni);
        uint32_t lo[PV], hi[PV];
        if (ok) {
            if (kuni != kd_key) {
                kd_key = kuni;
                kl = load_keydir(ix.d_mlast.kd, kuni);
                kf = load_keydir(ix.d_mfirst.kd, kuni);
            }
            // merged intervals overlapping [qs,qe] are exactly [first j with last_j >= qs, first j with first_j > qe)
            const bool mono_s = chunk_mono(t.qs, t.full), mono_e = chunk_mono(t.qe, t.full);
            if (!(mono_s && mono_e &&
                  warp_window_ranks2<true, false>(cl, ix.d_mlast, kl, t.qs, t.s_qs, aux0, lo, cf, ix.d_mfirst, kf, t.qe, t.s_qe, aux1,
                                                  hi, kuni))) {
                ok = warp_window_ranks<true>(cl, ix.d_mlast, kl, kuni, t.qs, t.active, mono_s, t.s_qs, aux0, lo);
                ok = ok && warp_ranks_le_end(cf, ix.d_mfirst, kf, kuni, t.qs, t.qe, t.active, mono_s, mono_e, t.s_qs, t.s_qe, aux1, hi);
            }
        }
        if (ok) {
            uint32_t c32[PV];
#if BRG_COV_COMPACT
            // Sparse target sets: most reads touch nothing and get 0.  The rows that do overlap something are compacted
            // across the warp and only they run the epilogue -- 2 passes of 32 rows instead of 8 for a typical chunk of the
            // 100M x 1M workload (16 % of the rows overlap).  Rows are listed in aux1 / aux2, results land in the staging area
            // of the transposed store, the query values come from the padded tile.
            uint32_t n_mine = 0;
#pragma unroll
            for (int j = 0; j < PV; j++) n_mine += (lo[j] < hi[j]) ? 1u : 0u;
            const uint32_t n_list = __reduce_add_sync(0xFFFFFFFFu, n_mine);
            if (t.full && n_list <= (uint32_t)COV_LIST_CAP) {
                const int lane = threadIdx.x & 31;
                __syncwarp();
                uint4* x4 = hist_lane_bins(aux0, lane);
                x4[0] = make_uint4(0, 0, 0, 0);
                x4[1] = make_uint4(0, 0, 0, 0);
                uint32_t at = n_mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, at, o);
                    if (lane >= o) at += y;
                }
                at -= n_mine;
                uint32_t* L = aux1;  // [row | lo | hi], COV_LIST_CAP entries each (aux1 and aux2 are contiguous)
#pragma unroll
                for (int j = 0; j < PV; j++)
                    if (lo[j] < hi[j]) { L[at] = lane * PV + j; L[COV_LIST_CAP + at] = lo[j]; L[2 * COV_LIST_CAP + at] = hi[j]; at++; }
                __syncwarp();
                for (uint32_t e = lane; e < n_list; e += 32) {
                    const uint32_t slot = hist_slot(L[e]);
                    aux0[slot] = (uint32_t)coverage_from_range<false>(ix, L[COV_LIST_CAP + e], L[2 * COV_LIST_CAP + e], t.s_qs[slot], t.s_qe[slot]);
                }
                store_chunk_i64_staged32(aux0, out + (t.row0 - lane * PV), true);  // i32 accumulator, widened (interval_tree.rs:207)
                fails = 0;
                pipe.release();
                continue;
            }
#endif
#pragma unroll
            for (int j = 0; j < PV; j++) c32[j] = t.active ? (uint32_t)coverage_from_range<false>(ix, lo[j], hi[j], t.qs[j], t.qe[j]) : 0u;
            if (t.full) {
                store_chunk_i64_from32(aux0, out + (t.row0 - (threadIdx.x & 31) * PV), c32, true);  // i32 accumulator, widened (interval_tree.rs:207)
            } else if (t.active) {
                long long cov[PV];
#pragma unroll
                for (int j = 0; j < PV; j++) cov[j] = (long long)(int32_t)c32[j];
                store8_i64(out + t.row0, cov);
            }
            fails = 0;
        } else {
            cl.valid = false;
            cf.valid = false;
            todo_push(todo, t.chunk);
            fails++;
        }
        if (fails >= TP_BAIL || cl.wide || cf.wide) { pipe.bail_out(todo); break; }
        pipe.release();
    }
}

__global__ void __launch_bounds__(TP_WTILE) coverage_kernel_todo(IndexView ix, const uint32_t* __restrict__ key,
                                                                 const int32_t* __restrict__ start,
                                                                 const int32_t* __restrict__ end, uint64_t n_rows,
                                                                 int strict, int64_t* __restrict__ out, TodoList todo) {
    const uint32_t n_todo = *todo.count;
    if (blockIdx.x >= n_todo) return;  // (nothing was deferred: skip staging the key constants)
    __shared__ KeyDir skeys_buf[CI_SMEM_KEYS];
    const KeyDir* skeys = stage_coverage_keys(ix, skeys_buf);
    for (uint32_t c = blockIdx.x; c < n_todo; c += gridDim.x) {
        const uint64_t i = (uint64_t)todo.chunks[c] * TP_WTILE + threadIdx.x;
        if (i >= n_rows) continue;
        int32_t qs = ld_stream(start + i), qe = ld_stream(end + i);
        const uint32_t k = load_key(key, i);
        if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); }
        int32_t acc = 0;
        if (k < ix.n_keys) {
            uint32_t lo, hi;
            coverage_row_range(ix, skeys, k, qs, qe, &lo, &hi);
            acc = coverage_from_range<true>(ix, lo, hi, qs, qe);
        }
        st_stream(out + i, (int64_t)acc);
    }
}

// ---------------------------------------------------------------------------------------------------
// nearest
// ---------------------------------------------------------------------------------------------------
struct Cand {
    int32_t s, e;
    uint32_t pos;
};
// nearest_index.rs:252-260 (same precedence of the two tests, written as selects: no divergent branches).  Both
// differences are taken where they are positive, so they are exact in 32-bit unsigned arithmetic (< 2^32).
__device__ __forceinline__ uint32_t cand_distance_u32(int32_t qs, int32_t qe, int32_t is, int32_t ie) {
    const uint32_t right = (uint32_t)is - (uint32_t)qe, left = (uint32_t)qs - (uint32_t)ie;
    const uint32_t d = (ie < qs) ? left : 0u;
    return (qe < is) ? right : d;
}
__device__ __forceinline__ int64_t cand_distance(int32_t qs, int32_t qe, int32_t is, int32_t ie) {
    return (int64_t)cand_distance_u32(qs, qe, is, ie);
}
// cmp_candidate(a,b) <= 0  (nearest_index.rs:262-266): (distance, start, end, position)
__device__ __forceinline__ bool cand_le(int32_t qs, int32_t qe, const Cand& a, const Cand& b) {
    const uint32_t ad = cand_distance_u32(qs, qe, a.s, a.e), bd = cand_distance_u32(qs, qe, b.s, b.e);
    const bool tail = (a.e != b.e) ? a.e < b.e : a.pos <= b.pos;
    const bool mid = (a.s != b.s) ? a.s < b.s : tail;
    return (ad != bd) ? ad < bd : mid;
}
__device__ __forceinline__ Cand load_s(const IndexView& ix, uint32_t i) {
    const int4 r = __ldg(ix.s_rec + i);
    return Cand{r.x, r.y, (uint32_t)r.z};
}
__device__ __forceinline__ Cand load_e(const IndexView& ix, uint32_t i) {
    const int4 r = __ldg(ix.e_rec + i);
    return Cand{r.x, r.y, (uint32_t)r.z};
}

// one probe: returns true and fills `best` when an interval is found
// nearest_one through the nearest bins (unsorted probe tables): ONE 32-byte record gives n_le = #(start <= qe) and
// i_pme = first row whose prefix_max_end >= qs, then one or two 16-byte candidate records -- 3 divergent loads per row where
// the rank directories below take about a dozen.  The index holds no row with start > end (the bins are not built otherwise),
// so without an overlap every row that starts by qe has ended before qs and every later row ends after qe: #(end < qs) = n_le.
__device__ __forceinline__ bool nearest_one_bins(const IndexView& ix, const KeyDir& k, uint32_t c, int32_t qs, int32_t qe,
                                                 int include_overlaps, Cand* best) {
    if (k.seg_begin == k.seg_end) return false;                         // is_empty (nearest_index.rs:92-94)
    uint32_t n_le, i_pme;
    const uint32_t missing = bins_pair_ranks(ix.nb, k, qe, qs, &n_le, &i_pme);
    if (missing & 1u) n_le = rank_le_kd(ix.d_sstart, load_keydir(ix.d_sstart.kd, c), qe);
    if (missing & 2u) i_pme = rank_lt_kd(ix.d_pme, load_keydir(ix.d_pme.kd, c), qs);
    const bool proper = !(qe < qs);
    if (include_overlaps && proper && i_pme < n_le) {                   // first_overlap_by_start (:222-234)
        *best = load_s(ix, i_pme);
        return true;
    }
    if (include_overlaps == 2) return false;                            // overlaps only (BRG_NEAREST_OVERLAP_ONLY)
    // nearest_non_overlap_one (:192-220): left = last row in (end, start, position) order with end < qs
    uint32_t left_idx = n_le;
    if (!(include_overlaps && proper))                                  // (overlaps skipped or a degenerate probe: rank the ends)
        left_idx = rank_lt_kd(ix.d_eend, load_keydir(ix.d_eend.kd, c), qs);
    const bool has_left = left_idx > k.seg_begin, has_right = n_le < k.seg_end;
    if (!has_left && !has_right) return false;
    Cand l, r;
    if (has_left) l = load_e(ix, left_idx - 1);
    if (has_right) r = load_s(ix, n_le);
    if (has_left && has_right) *best = cand_le(qs, qe, l, r) ? l : r;
    else *best = has_left ? l : r;
    return true;
}
// the nearest bins' per-key constants in shared memory (all threads of the CTA; nullptr when absent or too many keys)
constexpr uint32_t NB_SMEM_KEYS = 64;
__device__ __forceinline__ const KeyDir* stage_nearest_keys(const IndexView& ix, KeyDir* smem) {
    if (!ix.nb.rec || ix.n_keys > NB_SMEM_KEYS) return nullptr;
    const uint32_t words = ix.n_keys * (uint32_t)(sizeof(KeyDir) / 4);
    const uint32_t* g = reinterpret_cast<const uint32_t*>(ix.nb.kd);
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem);
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) sm[i] = __ldg(g + i);
    __syncthreads();
    return smem;
}
__device__ __forceinline__ bool nearest_one(const IndexView& ix, uint32_t c, int32_t qs, int32_t qe,
                                            int include_overlaps, Cand* best, const KeyDir* skeys = nullptr) {
    if (ix.nb.rec) return nearest_one_bins(ix, skeys ? skeys[c] : load_keydir(ix.nb.kd, c), c, qs, qe, include_overlaps, best);
    KeyDir ks = load_keydir(ix.d_sstart.kd, c);
    if (ks.seg_begin == ks.seg_end) return false;                       // is_empty (nearest_index.rs:92-94)
    const uint32_t n_le = rank_le_kd(ix.d_sstart, ks, qe);              // partition_point(first <= end)
    if (include_overlaps && !(qe < qs) && n_le > ks.seg_begin && __ldg(ix.s_rec + n_le - 1).w >= qs) {
        KeyDir kp = load_keydir(ix.d_pme.kd, c);                        // first_overlap_by_start (:222-234)
        uint32_t idx = rank_lt_kd(ix.d_pme, kp, qs);                    // first i with prefix_max_end[i] >= qs
        *best = load_s(ix, idx);
        return true;
    }
    if (include_overlaps == 2) return false;                            // overlaps only (BRG_NEAREST_OVERLAP_ONLY)
    KeyDir ke = load_keydir(ix.d_eend.kd, c);                           // nearest_non_overlap_one (:192-220)
    const uint32_t left_idx = rank_lt_kd(ix.d_eend, ke, qs);            // partition_point(last < start)
    const bool has_left = left_idx > ke.seg_begin, has_right = n_le < ks.seg_end;
    if (!has_left && !has_right) return false;
    Cand l, r;
    if (has_left) l = load_e(ix, left_idx - 1);
    if (has_right) r = load_s(ix, n_le);
    if (has_left && has_right) *best = cand_le(qs, qe, l, r) ? l : r;
    else *best = has_left ? l : r;
    return true;
}

__global__ void __launch_bounds__(PB_THREADS) nearest_k1_kernel(IndexView ix, const uint32_t* __restrict__ key,
                                                                const int32_t* __restrict__ start,
                                                                const int32_t* __restrict__ end, uint64_t n,
                                                                int include_overlaps, int strict,
                                                                uint32_t* __restrict__ left_idx,
                                                                uint32_t* __restrict__ valid_bits,
                                                                int64_t* __restrict__ dist) {
    __shared__ KeyDir skeys_buf[NB_SMEM_KEYS];
    const KeyDir* skeys = stage_nearest_keys(ix, skeys_buf);
    // whole warps stay converged so that the ballot below packs 32 validity bits per store
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (n + 31) & ~31ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool found = false;
        Cand b{0, 0, 0};
        int32_t rs = 0, re = 0;
        if (i < n) {
            rs = ld_stream(start + i); re = ld_stream(end + i);
            uint32_t c = load_key(key, i);
            int32_t qs = rs, qe = re;
            if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); }  // nearest.rs:341-344
            if (c < ix.n_keys) found = nearest_one(ix, c, qs, qe, include_overlaps, &b, skeys);
        }
        uint32_t bits = __ballot_sync(0xFFFFFFFFu, found);
        if ((threadIdx.x & 31) == 0) valid_bits[i >> 5] = bits;  // Arrow validity bitmap, LSB first
        if (i < n) {
            st_stream(left_idx + i, found ? b.pos : 0u);
            // distance from RAW probe coordinates (nearest.rs:367-375)
            if (dist) st_stream(dist + i, found ? cand_distance(rs, re, b.s, b.e) : (int64_t)0);
        }
    }
}

// nearest_one with the three ranks already known:
//   n_le = seg_begin + #(starts <= qe), i_pme = first i with prefix_max_end[i] >= qs, left_idx = seg_begin + #(ends < qs)
// Branch-free: the two candidate records are loaded unconditionally from clamped indices, so the loads of the PV rows a
// lane owns are all in flight together; the choice is made with selects.  Requires seg_begin < seg_end.
__device__ __forceinline__ bool nearest_from_ranks(const IndexView& ix, uint32_t seg_begin, uint32_t seg_end, uint32_t n_le,
                                                   uint32_t i_pme, uint32_t left_idx, int32_t qs, int32_t qe,
                                                   int include_overlaps, Cand* best) {
    const uint32_t last = seg_end - 1;
    const bool has_left = left_idx > seg_begin, has_right = n_le < seg_end;
    // first_overlap_by_start (:222-234): there is an overlap iff prefix_max_end[n_le - 1] >= qs, and prefix_max_end is
    // non-decreasing: iff the first row whose prefix_max_end >= qs (= i_pme, already known) lies before n_le.  No load needed
    // for the test -- and the overlap record and the right neighbour are never needed together, so one load serves both.
    const bool ov = include_overlaps && !(qe < qs) && i_pme < n_le;
    const int4 rs = __ldg(ix.s_rec + min(ov ? i_pme : n_le, last));           // the overlapping row, else the right neighbour (first start > qe)
    const int4 rl = __ldg(ix.e_rec + (has_left ? left_idx - 1 : seg_begin));  // left neighbour: last end < qs
    const Cand l{rl.x, rl.y, (uint32_t)rl.z}, r{rs.x, rs.y, (uint32_t)rs.z};
    const bool take_left = !ov && has_left && (!has_right || cand_le(qs, qe, l, r));
    *best = take_left ? l : r;
    return ov || (include_overlaps != 2 && (has_left || has_right));  // 2 = overlaps only
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2) nearest_k1_kernel_tiled(IndexView ix, const uint32_t* __restrict__ key,
                                                                         const int32_t* __restrict__ start,
                                                                         const int32_t* __restrict__ end, uint64_t n_rows,
                                                                         int include_overlaps, int strict,
                                                                         uint32_t* __restrict__ left_idx,
                                                                         uint8_t* __restrict__ valid_bytes,
                                                                         int64_t* __restrict__ dist, TodoList todo) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpRing* rings = reinterpret_cast<WarpRing*>(smem_raw);  // one per warp of the CTA
    uint32_t* aux0 = rings[threadIdx.x >> 5].aux[0];
    uint32_t* aux1 = rings[threadIdx.x >> 5].aux[1];
    uint32_t* aux2 = rings[threadIdx.x >> 5].aux[2];
    ProbePipe pipe(rings, key, start, end, n_rows);  // n_rows is a multiple of 256 here: every chunk is full
    ProbeTile t;
    // cs is served by the single search every chunk (starts keyed on unsorted ends), cp / ce whenever one of their windows
    // spans more than one round and the interleaved dual search steps aside: far prefetch on all three, 0.737 -> 0.700 ms
    WarpCursorFar cs, cp, ce;
    cs.reset();
    cp.reset();
    ce.reset();
    KeyDir ks, kp, ke;
    uint32_t kd_key = 0xFFFFFFFFu;
    uint32_t fails = 0;  // consecutive deferred chunks
    while (pipe.next(t)) {
        if (strict) strict_adjust(t);
        pad_tile(t);
        uint32_t kuni;
        bool ok = warp_uniform_key(t, ix.n_keys, &kuni);
        uint32_t n_le[PV], i_pme[PV], l_idx[PV];
        if (ok) {
            if (kuni != kd_key) {
                kd_key = kuni;
                ks = load_keydir(ix.d_sstart.kd, kuni);
                kp = load_keydir(ix.d_pme.kd, kuni);
                ke = load_keydir(ix.d_eend.kd, kuni);
            }
            const bool mono_s = chunk_mono(t.qs, t.full), mono_e = chunk_mono(t.qe, t.full);
            // the two start-keyed searches interleaved when both windows are short, else one after the other (forward searches
            // for these two as well lose: 0.60 vs 0.567 ms -- for sorted queries the inverse search is the cheaper one)
            const bool dual = mono_s && warp_window_ranks2<true, true>(cp, ix.d_pme, kp, t.qs, t.s_qs, aux1, i_pme, ce, ix.d_eend, ke,
                                                                       t.qs, t.s_qs, aux2, l_idx, kuni);
            ok = warp_ranks_le_end(cs, ix.d_sstart, ks, kuni, t.qs, t.qe, t.active, mono_s, mono_e, t.s_qs, t.s_qe, aux0, n_le);
            if (!dual) {
                ok = ok && warp_window_ranks<true>(cp, ix.d_pme, kp, kuni, t.qs, t.active, mono_s, t.s_qs, aux1, i_pme);
                ok = ok && warp_window_ranks<true>(ce, ix.d_eend, ke, kuni, t.qs, t.active, mono_s, t.s_qs, aux2, l_idx);
            }
        }
        if (ok) {
#if BRG_NEAREST_PREFETCH
            // The candidate records of the NEXT chunk sit where this chunk's windows ended (the cursors): s_rec / e_rec are
            // 160 MB for 5M build rows, more than the L2 keeps next to the probe stream, so the first touch of each record
            // would wait for DRAM in the middle of the next chunk's epilogue.
            {
                const uint32_t lane_ = threadIdx.x & 31, last_ = ks.seg_end ? ks.seg_end - 1 : 0u;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ix.s_rec + min(cs.pos + lane_, last_)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ix.e_rec + min(ce.pos + lane_, last_)));
            }
#endif
            uint32_t pos[PV];
            long long dd[PV];
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < PV; j++) {
                Cand b{0, 0, 0};
                const bool found = ks.seg_begin != ks.seg_end &&  // is_empty (nearest_index.rs:92-94); warp-uniform
                                   nearest_from_ranks(ix, ks.seg_begin, ks.seg_end, n_le[j], i_pme[j], l_idx[j], t.qs[j], t.qe[j],
                                                      include_overlaps, &b);
                bits |= found ? (1u << j) : 0u;
                pos[j] = found ? b.pos : 0u;
                // distance from RAW probe coordinates (nearest.rs:367-375): undo the strict adjustment
                const int32_t rs = strict ? wsub(t.qs[j], 1) : t.qs[j], re = strict ? wadd(t.qe[j], 1) : t.qe[j];
                dd[j] = found ? cand_distance(rs, re, b.s, b.e) : 0ll;
            }
            // (transposing these stores through shared memory, as count / coverage do, changes nothing here: 0.767 vs 0.772 ms --
            // the kernel is bound by its ~1900 warp instructions per chunk, not by the store path)
            uint4* lp = reinterpret_cast<uint4*>(left_idx + t.row0);
            __stcs(lp, make_uint4(pos[0], pos[1], pos[2], pos[3]));
            __stcs(lp + 1, make_uint4(pos[4], pos[5], pos[6], pos[7]));
            valid_bytes[t.row0 >> 3] = (uint8_t)bits;  // Arrow validity bitmap, LSB first: byte r/8 holds rows r..r+7
            if (dist) store8_i64(dist + t.row0, dd);
            fails = 0;
        } else {
            cs.valid = false;
            cp.valid = false;
            ce.valid = false;
            todo_push(todo, t.chunk);
            fails++;
        }
        if (fails >= TP_BAIL || cs.wide || cp.wide || ce.wide) { pipe.bail_out(todo); break; }
        pipe.release();
    }
}

// ---------------------------------------------------------------------------------------------------
// nearest, k = 1, WINDOWED (round 2): the window machinery of the one-launch pair join (brg_win.cuh) instead of per-warp cursors
// and inverse searches.  A warp takes 256-row chunks from a contiguous range; a chunk is two halves of 128 rows, lane l owning
// rows l, l + 32, l + 64, l + 96 of a half (32 consecutive sorted queries per search step: broadcasts and consecutive banks;
// every output store of the warp writes 128 / 256 contiguous bytes; a ballot is one word of the validity bitmap).  Per half:
//   * two windows by TMA, each starting at a directory bin (one dependent load): S-order rows from the bin of qs_first - 1 in the
//     prefix_max_end directory (start, prefix_max_end and the 16-byte records {start, end, row id, prefix_max_end}), E-order rows
//     from one row before the bin of qs_first - 1 in the end directory (end and the records {start, end, row id});
//   * the three ranks of nearest_one (nearest_index.rs:91-101,192-234) -- #(start <= qe), first prefix_max_end >= qs, #(end < qs) --
//     by branch-free searches over the windows; the two candidate records per row come from shared memory.
// 64 registers / 32 warps per SM instead of 128 / 16.  (An L2 prefetch of the warp's next chunk of the probe stream costs more issue
// slots than it saves latency: 0.458 vs 0.441 ms.)  A half that does not fit (key change, unsorted starts, end < start, window
// longer than NW_WCAP rows) sends its chunk to the one-row-per-thread kernel through the to-do list; exact for any input.
// ---------------------------------------------------------------------------------------------------
constexpr int NW_PV = 4;
constexpr int NW_ROWS = 32 * NW_PV;
#ifndef BRG_NW_WCAP
#define BRG_NW_WCAP 48
#endif
#ifndef BRG_NW_WARPS
#define BRG_NW_WARPS 8
#endif
#ifndef BRG_NW_CTAS
#define BRG_NW_CTAS 4
#endif
constexpr int NW_WCAP = BRG_NW_WCAP, NW_PAD = win_pad(NW_WCAP), NW_WARPS = BRG_NW_WARPS;
static_assert(TP_WTILE == 2 * NW_ROWS, "a to-do chunk is two halves");
struct __align__(16) NearWarp {
    int32_t start_[NW_WCAP + 4 + NW_PAD];  // window rows from [4] on; [3] is the guard word read at index -1; sentinels behind
    int32_t pme_[NW_WCAP + 4 + NW_PAD];
    int32_t eend_[NW_WCAP + 4 + NW_PAD];
    int4 s_rec[NW_WCAP];
    int4 e_rec[NW_WCAP];
    uint64_t bar[2];
};

__global__ void __launch_bounds__(NW_WARPS * 32, BRG_NW_CTAS)
nearest_k1_kernel_win(IndexView ix, const uint32_t* __restrict__ key, const int32_t* __restrict__ start,
                      const int32_t* __restrict__ end, uint64_t n_rows, int include_overlaps, int strict,
                      uint32_t* __restrict__ left_idx, uint32_t* __restrict__ valid_bits, int64_t* __restrict__ dist, TodoList todo) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    NearWarp& sm = reinterpret_cast<NearWarp*>(smem_raw)[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31;
    int32_t* const w_start = sm.start_ + 4;
    int32_t* const w_pme = sm.pme_ + 4;
    int32_t* const w_eend = sm.eend_ + 4;
    if (lane == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); mbar_fence_init(); }
    win_sentinels_init<NW_WCAP>(w_start, lane);  // the rows behind a full window: once per kernel
    win_sentinels_init<NW_WCAP>(w_pme, lane);
    win_sentinels_init<NW_WCAP>(w_eend, lane);
    __syncwarp();
    uint32_t ph_s = 0, ph_e = 0;
    const uint64_t n_chunks = n_rows / TP_WTILE;  // n_rows is a multiple of 256 here
    const uint64_t n_warps = (uint64_t)gridDim.x * NW_WARPS, wid = (uint64_t)blockIdx.x * NW_WARPS + (threadIdx.x >> 5);
    const uint64_t c_begin = n_chunks * wid / n_warps, c_end = n_chunks * (wid + 1) / n_warps;
    uint32_t fails = 0;
    const int32_t sdelta = strict ? 1 : 0;
    for (uint64_t chunk = c_begin; chunk < c_end; chunk++) {
        bool ok = true;
        for (int half = 0; half < 2 && ok; half++) {
            const uint64_t row0 = chunk * TP_WTILE + (uint64_t)half * NW_ROWS;
            int32_t rs[NW_PV], re[NW_PV], qs[NW_PV], qe[NW_PV];
            uint32_t kk[NW_PV];
#pragma unroll
            for (int j = 0; j < NW_PV; j++) {
                const uint64_t r = row0 + lane + 32u * j;
                rs[j] = __ldcs(start + r);
                re[j] = __ldcs(end + r);
                kk[j] = key ? __ldcs(key + r) : 0u;
                qs[j] = wadd(rs[j], sdelta);  // nearest.rs:341-344
                qe[j] = wsub(re[j], sdelta);
            }
            const uint32_t k0 = __shfl_sync(0xFFFFFFFFu, kk[0], 0);
            bool fits = true;
#pragma unroll
            for (int j = 0; j < NW_PV; j++) {
                int32_t nx = __shfl_down_sync(0xFFFFFFFFu, qs[j], 1);
                if (j + 1 < NW_PV) {
                    const int32_t wrap = __shfl_sync(0xFFFFFFFFu, qs[j + 1 < NW_PV ? j + 1 : j], 0);
                    if (lane == 31) nx = wrap;
                } else if (lane == 31) {
                    nx = INT32_MAX;
                }
                fits = fits & (kk[j] == k0) & (qs[j] <= nx) & !(qe[j] < qs[j]);
            }
            if (!__all_sync(0xFFFFFFFFu, fits)) { ok = false; break; }
            uint32_t pos[NW_PV];
            long long dd[NW_PV];
            bool found[NW_PV];
#pragma unroll
            for (int j = 0; j < NW_PV; j++) { pos[j] = 0u; dd[j] = 0ll; found[j] = false; }
            // the key's segment and the two window starts: lane 0 the S-order window, lane 1 the E-order one
            uint32_t wbeg = 0, seg_b = 0, seg_e = 0;
            if (k0 < ix.n_keys && lane < 2) {
                const DirView dv = lane ? ix.d_eend : ix.d_pme;
                const KeyDir k = load_keydir(dv.kd, k0);
                wbeg = win_begin(dv, k, __shfl_sync(3u, qs[0], 0));
                seg_b = k.seg_begin; seg_e = k.seg_end;
            }
            const uint32_t Ws = __shfl_sync(0xFFFFFFFFu, wbeg, 0), We = __shfl_sync(0xFFFFFFFFu, wbeg, 1);
            const uint32_t sb = __shfl_sync(0xFFFFFFFFu, seg_b, 0), se = __shfl_sync(0xFFFFFFFFu, seg_e, 0);  // (same rows in both orders)
            if (k0 < ix.n_keys && sb < se) {  // else: key absent or without build rows -- is_empty (nearest_index.rs:92-94): NULL rows
                // S-order window from Ws (at least the key's last row); E-order window from ONE ROW BEFORE We: the left neighbour of
                // the chunk's first row may be the last row in front of the bin
                const uint32_t Was = min(Ws, se - 1) & ~3u, Wae = min(We > sb ? We - 1 : We, se - 1) & ~3u;
                const uint32_t nws = min((uint32_t)NW_WCAP, se - Was), nwe = min((uint32_t)NW_WCAP, se - Wae);
                const uint32_t wlos = max(sb, Was) - Was, wloe = max(sb, Wae) - Wae;
                const bool trunc_s = Was + nws < se, trunc_e = Wae + nwe < se;
                __syncwarp();  // the previous half's readers are done with the windows
                if (lane == 0) {
                    const uint32_t bs = ((nws + 3u) & ~3u) * 4u, be = ((nwe + 3u) & ~3u) * 4u;
                    mbar_expect_tx(&sm.bar[0], 2u * bs + nws * 16u);
                    tma_load_1d(w_start, ix.s_start + Was, bs, &sm.bar[0], 0ull);
                    tma_load_1d(w_pme, ix.pme + Was, bs, &sm.bar[0], 0ull);
                    tma_load_1d(sm.s_rec, ix.s_rec + Was, nws * 16u, &sm.bar[0], 0ull);
                    mbar_expect_tx(&sm.bar[1], be + nwe * 16u);
                    tma_load_1d(w_eend, ix.e_end + Wae, be, &sm.bar[1], 0ull);
                    tma_load_1d(sm.e_rec, ix.e_rec + Wae, nwe * 16u, &sm.bar[1], 0ull);
                }
                mbar_wait(&sm.bar[0], ph_s);
                ph_s ^= 1u;
                mbar_wait(&sm.bar[1], ph_e);
                ph_e ^= 1u;
                if (nws < (uint32_t)NW_WCAP || nwe < (uint32_t)NW_WCAP) {  // (a window cut short by the end of the key; warp-uniform)
                    win_sentinels_short<NW_WCAP>(w_start, nws, lane);
                    win_sentinels_short<NW_WCAP>(w_pme, nws, lane);
                    win_sentinels_short<NW_WCAP>(w_eend, nwe, lane);
                    __syncwarp();
                }
                uint32_t n_le[NW_PV], i_pme[NW_PV], l_idx[NW_PV];
                win_ranks<true, NW_PV, NW_WCAP>(w_start, wlos, nws, qe, n_le);     // #(start <= qe)              partition_point(first <= end)
                win_ranks<false, NW_PV, NW_WCAP>(w_pme, wlos, nws, qs, i_pme);     // first prefix_max_end >= qs    first_overlap_by_start (:222-234)
                win_ranks<false, NW_PV, NW_WCAP>(w_eend, wloe, nwe, qs, l_idx);    // #(end < qs)                 partition_point(last < start)
                bool beyond = false;
#pragma unroll
                for (int j = 0; j < NW_PV; j++) beyond = beyond | (trunc_s & (n_le[j] == nws)) | (trunc_e & (l_idx[j] == nwe));
                if (__any_sync(0xFFFFFFFFu, beyond)) { ok = false; break; }
                // nearest_one with the ranks known (see nearest_from_ranks): the overlapping row and the right neighbour are never needed
                // together, so two records per row.  cmp_candidate (nearest_index.rs:262-266) orders by (distance, start, end, row id):
                // the tie-break beyond the distance is evaluated only when some row of the warp needs it.
#pragma unroll
                for (int j = 0; j < NW_PV; j++) {
                    const bool has_left = Wae + l_idx[j] > sb, has_right = Was + n_le[j] < se;
                    const bool ov = include_overlaps && i_pme[j] < n_le[j];
                    const int4 r4 = sm.s_rec[min(ov ? i_pme[j] : n_le[j], nws - 1u)];
                    const int4 l4 = sm.e_rec[has_left ? l_idx[j] - 1u : 0u];
                    const Cand l{l4.x, l4.y, (uint32_t)l4.z}, r{r4.x, r4.y, (uint32_t)r4.z};
                    const uint32_t dl = cand_distance_u32(qs[j], qe[j], l.s, l.e), dr = cand_distance_u32(qs[j], qe[j], r.s, r.e);
                    const bool need = !ov && has_left;         // the left neighbour is a candidate
                    bool left_le = !has_right || dl < dr;      // ... and wins without a tie-break
                    if (__any_sync(0xFFFFFFFFu, need & has_right & (dl == dr))) {
                        const bool tail = (l.e != r.e) ? l.e < r.e : l.pos <= r.pos;
                        const bool mid = (l.s != r.s) ? l.s < r.s : tail;
                        left_le = left_le | ((dl == dr) & mid);
                    }
                    const bool take_left = need & left_le;
                    const Cand b = take_left ? l : r;
                    found[j] = ov || (include_overlaps != 2 && (has_left || has_right));  // 2 = overlaps only
                    pos[j] = found[j] ? b.pos : 0u;
                    // distance from RAW probe coordinates (nearest.rs:367-375); without `strict` they are the searched ones
                    const uint32_t d = strict ? cand_distance_u32(rs[j], re[j], b.s, b.e) : (take_left ? dl : dr);
                    dd[j] = found[j] ? (long long)d : 0ll;
                }
            }
            uint32_t word = 0;  // lane j < PV ends up with the validity word of rows 32 j .. 32 j + 31 (Arrow bitmap, LSB first)
#pragma unroll
            for (int j = 0; j < NW_PV; j++) {
                const uint64_t r = row0 + lane + 32u * j;
                const uint32_t bits = __ballot_sync(0xFFFFFFFFu, found[j]);
                word = lane == (uint32_t)j ? bits : word;
                __stcs(left_idx + r, pos[j]);
                if (dist) __stcs(reinterpret_cast<long long*>(dist) + r, dd[j]);
            }
            if (lane < (uint32_t)NW_PV) valid_bits[(row0 >> 5) + lane] = word;
        }
        if (ok) {
            fails = 0;
        } else {
            todo_push(todo, chunk);
            if (++fails >= TP_BAIL) {  // this range does not look like a sorted table over a sparse build side: hand the rest over
                for (uint64_t c = chunk + 1 + lane; c < c_end; c += 32) todo.chunks[atomicAdd(todo.count, 1u)] = (uint32_t)c;
                break;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// count_overlaps and coverage, WINDOWED (round 2; same scheme as nearest_k1_kernel_win): two ranks per row from branch-free
// searches over TMA windows, lane l owning rows l + 32 j of a 128-row half chunk.
//   count    (interval_tree.rs:41-49):  #(start <= qe) over an S-order window  -  #(end < qs) over an E-order window
//   coverage (interval_tree.rs:145-152): [first merged interval with last >= qs, first with first > qe) over ONE window of the
//            merged intervals (first, last and the records {first, last, pre[j], pre[j+1]}), then the clipped-boundary formula
// ---------------------------------------------------------------------------------------------------
#ifndef BRG_COUNT_WIN_DEFAULT
#define BRG_COUNT_WIN_DEFAULT false
#endif
#ifndef BRG_COV_WIN_DEFAULT
#define BRG_COV_WIN_DEFAULT false
#endif
#ifndef BRG_R2_WCAP
#define BRG_R2_WCAP 28   // bit_floor = 16: five search steps cover up to 31 rows
#endif
#ifndef BRG_R2_WARPS
#define BRG_R2_WARPS 8
#endif
#ifndef BRG_R2_CTAS
#define BRG_R2_CTAS 4
#endif
constexpr int R2_WCAP = BRG_R2_WCAP, R2_PAD = win_pad(R2_WCAP), R2_WARPS = BRG_R2_WARPS;
struct __align__(16) Rank2Warp {
    int32_t a_[R2_WCAP + 4 + R2_PAD];  // searched with `<= qe`: starts (count) / merged firsts (coverage); rows from [4] on
    int32_t b_[R2_WCAP + 4 + R2_PAD];  // searched with `< qs`:  ends (count) / merged lasts (coverage)
    int4 rec[R2_WCAP];                 // coverage: m_rec of the window
    uint64_t bar[2];
};

template <int OP>  // 0 = count_overlaps, 1 = coverage
__global__ void __launch_bounds__(R2_WARPS * 32, BRG_R2_CTAS)
rank2_kernel_win(IndexView ix, const uint32_t* __restrict__ key, const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                 uint64_t n_rows, int strict, int64_t* __restrict__ out, TodoList todo) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Rank2Warp& sm = reinterpret_cast<Rank2Warp*>(smem_raw)[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31;
    int32_t* const w_a = sm.a_ + 4;
    int32_t* const w_b = sm.b_ + 4;
    if (lane == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); mbar_fence_init(); }
    win_sentinels_init<R2_WCAP>(w_a, lane);
    win_sentinels_init<R2_WCAP>(w_b, lane);
    __syncwarp();
    uint32_t ph_a = 0, ph_b = 0;
    const DirView& da = OP == 0 ? ix.d_sstart : ix.d_mlast;  // the directory that places window A (coverage: one window, placed by
    const DirView& db = OP == 0 ? ix.d_eend : ix.d_mlast;    // the lasts: an interval that ends before the chunk also starts before it)
    const int32_t* ga = OP == 0 ? ix.s_start : ix.m_first;
    const int32_t* gb = OP == 0 ? ix.e_end : ix.m_last;
    const uint64_t n_chunks = n_rows / TP_WTILE;  // n_rows is a multiple of 256 here
    const uint64_t n_warps = (uint64_t)gridDim.x * R2_WARPS, wid = (uint64_t)blockIdx.x * R2_WARPS + (threadIdx.x >> 5);
    const uint64_t c_begin = n_chunks * wid / n_warps, c_end = n_chunks * (wid + 1) / n_warps;
    const int32_t sdelta = strict ? 1 : 0;
    uint32_t fails = 0;
    for (uint64_t chunk = c_begin; chunk < c_end; chunk++) {
        bool ok = true;
        for (int half = 0; half < 2 && ok; half++) {
            const uint64_t row0 = chunk * TP_WTILE + (uint64_t)half * NW_ROWS;
            int32_t qs[NW_PV], qe[NW_PV];
            uint32_t kk[NW_PV];
#pragma unroll
            for (int j = 0; j < NW_PV; j++) {
                const uint64_t r = row0 + lane + 32u * j;
                qs[j] = wadd(__ldcs(start + r), sdelta);
                qe[j] = wsub(__ldcs(end + r), sdelta);
                kk[j] = key ? __ldcs(key + r) : 0u;
            }
            const uint32_t k0 = __shfl_sync(0xFFFFFFFFu, kk[0], 0);
            bool fits = true;
#pragma unroll
            for (int j = 0; j < NW_PV; j++) {
                int32_t nx = __shfl_down_sync(0xFFFFFFFFu, qs[j], 1);
                if (j + 1 < NW_PV) {
                    const int32_t wrap = __shfl_sync(0xFFFFFFFFu, qs[j + 1 < NW_PV ? j + 1 : j], 0);
                    if (lane == 31) nx = wrap;
                } else if (lane == 31) {
                    nx = INT32_MAX;
                }
                fits = fits & (kk[j] == k0) & (qs[j] <= nx) & !(qe[j] < qs[j]);
                if (OP == 1) fits = fits & (qs[j] != INT32_MIN) & (qe[j] != INT32_MAX);  // (wrapping `start - 1` / `end + 1`: per-row path)
            }
            if (!__all_sync(0xFFFFFFFFu, fits)) { ok = false; break; }
            long long res[NW_PV];
#pragma unroll
            for (int j = 0; j < NW_PV; j++) res[j] = 0ll;
            const int32_t qs_first = __shfl_sync(0xFFFFFFFFu, qs[0], 0);
            uint32_t wbeg = 0, seg_b = 0, seg_e = 0;
            if (k0 < ix.n_keys && lane < 2) {
                const DirView dv = lane ? db : da;
                const KeyDir k = load_keydir(dv.kd, k0);
                wbeg = win_begin(dv, k, qs_first);
                seg_b = k.seg_begin; seg_e = k.seg_end;
            }
            const uint32_t Wa0 = __shfl_sync(0xFFFFFFFFu, wbeg, 0), Wb0 = __shfl_sync(0xFFFFFFFFu, wbeg, 1);
            const uint32_t sb = __shfl_sync(0xFFFFFFFFu, seg_b, 0), se = __shfl_sync(0xFFFFFFFFu, seg_e, 0);
            if (k0 < ix.n_keys && sb < se) {  // else: key absent or without build rows -> 0 (interval_tree.rs:265)
                const uint32_t Wa = min(Wa0, se - 1) & ~3u, Wb = min(Wb0, se - 1) & ~3u;
                const uint32_t nwa = min((uint32_t)R2_WCAP, se - Wa), nwb = min((uint32_t)R2_WCAP, se - Wb);
                const uint32_t wloa = max(sb, Wa) - Wa, wlob = max(sb, Wb) - Wb;
                const bool trunc_a = Wa + nwa < se, trunc_b = Wb + nwb < se;
                __syncwarp();  // the previous half's readers are done with the windows
                if (lane == 0) {
                    const uint32_t ba = ((nwa + 3u) & ~3u) * 4u, bb = ((nwb + 3u) & ~3u) * 4u;
                    mbar_expect_tx(&sm.bar[0], ba + (OP == 1 ? nwa * 16u : 0u));
                    tma_load_1d(w_a, ga + Wa, ba, &sm.bar[0], 0ull);
                    if (OP == 1) tma_load_1d(sm.rec, ix.m_rec + Wa, nwa * 16u, &sm.bar[0], 0ull);
                    mbar_expect_tx(&sm.bar[1], bb);
                    tma_load_1d(w_b, gb + Wb, bb, &sm.bar[1], 0ull);
                }
                mbar_wait(&sm.bar[0], ph_a);
                ph_a ^= 1u;
                mbar_wait(&sm.bar[1], ph_b);
                ph_b ^= 1u;
                if (nwa < (uint32_t)R2_WCAP || nwb < (uint32_t)R2_WCAP) {
                    win_sentinels_short<R2_WCAP>(w_a, nwa, lane);
                    win_sentinels_short<R2_WCAP>(w_b, nwb, lane);
                    __syncwarp();
                }
                uint32_t ra[NW_PV], rb[NW_PV];
                win_ranks<true, NW_PV, R2_WCAP>(w_a, wloa, nwa, qe, ra);   // #(start <= qe)   / first merged interval with first > qe
                win_ranks<false, NW_PV, R2_WCAP>(w_b, wlob, nwb, qs, rb);  // #(end < qs)      / first merged interval with last >= qs
                bool beyond = false;
#pragma unroll
                for (int j = 0; j < NW_PV; j++) beyond = beyond | (trunc_a & (ra[j] == nwa)) | (OP == 0 ? (trunc_b & (rb[j] == nwb)) : false);
                if (__any_sync(0xFFFFFFFFu, beyond)) { ok = false; break; }
#pragma unroll
                for (int j = 0; j < NW_PV; j++) {
                    if (OP == 0) {
                        res[j] = (long long)(Wa + ra[j]) - (long long)(Wb + rb[j]);  // (segments of the two orders coincide: global indices subtract)
                    } else {
                        // merged intervals [lo, hi) overlap the row; both indices relative to window A (= window B's rows: Wa == Wb)
                        const uint32_t lo = rb[j], hi = ra[j];
                        const int4 a = sm.rec[min(lo, nwa - 1u)], b = sm.rec[min(hi - 1u, nwa - 1u)];
                        int32_t acc = cov_term(qs[j], qe[j], a.x, a.y);
                        const int32_t rest = wadd(cov_term(qs[j], qe[j], b.x, b.y), wsub(b.z, a.w));
                        acc = (hi > lo + 1) ? wadd(acc, rest) : acc;
                        res[j] = (lo < hi) ? (long long)acc : 0ll;  // i32 accumulator, widened (interval_tree.rs:207)
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NW_PV; j++) __stcs(reinterpret_cast<long long*>(out) + row0 + lane + 32u * j, res[j]);
        }
        if (ok) {
            fails = 0;
        } else {
            todo_push(todo, chunk);
            if (++fails >= TP_BAIL) {
                for (uint64_t c = chunk + 1 + lane; c < c_end; c += 32) todo.chunks[atomicAdd(todo.count, 1u)] = (uint32_t)c;
                break;
            }
        }
    }
}

__global__ void __launch_bounds__(TP_WTILE) nearest_k1_kernel_todo(IndexView ix, const uint32_t* __restrict__ key,
                                                                   const int32_t* __restrict__ start,
                                                                   const int32_t* __restrict__ end, uint64_t n_rows,
                                                                   int include_overlaps, int strict,
                                                                   uint32_t* __restrict__ left_idx,
                                                                   uint32_t* __restrict__ valid_bits,
                                                                   int64_t* __restrict__ dist, TodoList todo) {
    const uint32_t n_todo = *todo.count;
    if (blockIdx.x >= n_todo) return;  // (nothing was deferred to this CTA: skip staging the key constants)
    __shared__ KeyDir skeys_buf[NB_SMEM_KEYS];
    const KeyDir* skeys = stage_nearest_keys(ix, skeys_buf);
    for (uint32_t c = blockIdx.x; c < n_todo; c += gridDim.x) {  // whole 256-row chunks: warps stay converged
        const uint64_t i = (uint64_t)todo.chunks[c] * TP_WTILE + threadIdx.x;
        const int32_t rs = ld_stream(start + i), re = ld_stream(end + i);
        const uint32_t k = load_key(key, i);
        int32_t qs = rs, qe = re;
        if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); }
        Cand b{0, 0, 0};
        const bool found = k < ix.n_keys && nearest_one(ix, k, qs, qe, include_overlaps, &b, skeys);
        const uint32_t bits = __ballot_sync(0xFFFFFFFFu, found);
        if ((threadIdx.x & 31) == 0) valid_bits[i >> 5] = bits;
        st_stream(left_idx + i, found ? b.pos : 0u);
        if (dist) st_stream(dist + i, found ? cand_distance(rs, re, b.s, b.e) : (int64_t)0);
    }
}

// k > 1: shared walker for the count and emit passes
template <bool EMIT>
__device__ __forceinline__ uint32_t nearest_k_walk(const IndexView& ix, uint32_t c, int32_t qs, int32_t qe, int32_t rs,
                                                   int32_t re, uint32_t k, int include_overlaps, uint32_t probe_row,
                                                   uint64_t out_base, uint32_t* build_idx, uint32_t* valid,
                                                   uint32_t* probe_idx, int64_t* dist) {
    uint32_t n_out = 0;
    auto emit = [&](const Cand& x) {
        if (EMIT) {
            uint64_t o = out_base + n_out;
            build_idx[o] = x.pos;
            valid[o] = 1u;
            probe_idx[o] = probe_row;
            if (dist) dist[o] = cand_distance(rs, re, x.s, x.e);
        }
        n_out++;
    };
    if (k == 1) {
        // nearest_k(k == 1) delegates to nearest_one (nearest_index.rs:110-115), whose overlap step is skipped for a degenerate
        // probe (end < start after the strict adjustment, :222-225) -- the walk below would report a spanning interval instead
        Cand b{0, 0, 0};
        if (nearest_one(ix, c, qs, qe, include_overlaps ? 1 : 0, &b)) emit(b);
        return n_out;
    }
    KeyDir ks = load_keydir(ix.d_sstart.kd, c);
    if (ks.seg_begin == ks.seg_end || k == 0) return 0;
    const uint32_t n_le = rank_le_kd(ix.d_sstart, ks, qe);
    if (include_overlaps && n_le > ks.seg_begin) {
        // tree.query + sort by (start,end,position) (nearest_index.rs:124-139): S-order already is that order;
        // rows before the first prefix_max_end >= qs cannot reach qs
        KeyDir kp = load_keydir(ix.d_pme.kd, c);
        uint32_t i0 = rank_lt_kd(ix.d_pme, kp, qs);
        for (uint32_t i = i0; i < n_le && n_out < k; i++) {
            const int4 r = __ldg(ix.s_rec + i);
            if (r.y >= qs) emit(Cand{r.x, r.y, (uint32_t)r.z});
        }
    }
    if (n_out == k) return n_out;
    KeyDir ke = load_keydir(ix.d_eend.kd, c);
    uint32_t left_idx = rank_lt_kd(ix.d_eend, ke, qs);
    uint32_t right_idx = n_le;
    // An interval can sit in BOTH lists (end < qs and start > qe) only for a degenerate probe (qe < qs) or an
    // inverted build interval; the reference de-duplicates with a hash set (`seen`, nearest_index.rs:121,186).
    // Here membership is decided by order: the left list is consumed downwards in (end,start,position) order and
    // the right list upwards in (start,end,position) order.
    while (n_out < k) {
        const bool has_left = left_idx > ke.seg_begin, has_right = right_idx < ks.seg_end;
        if (!has_left && !has_right) break;
        Cand l, r, next;
        if (has_left) l = load_e(ix, left_idx - 1);
        if (has_right) r = load_s(ix, right_idx);
        bool take_left = has_left && (!has_right || cand_le(qs, qe, l, r));
        bool dup = false;
        if (take_left) {
            left_idx--;
            next = l;
            // `seen` (nearest_index.rs:186): already emitted from the right list iff it starts after qe and sorts
            // before the current right head in (start,end,position) order
            if (l.s > qe) {
                if (!has_right) dup = true;
                else dup = (l.s != r.s) ? l.s < r.s : (l.e != r.e) ? l.e < r.e : l.pos < r.pos;
            }
        } else {
            right_idx++;
            next = r;
            if (r.e < qs) {
                // already emitted from the left list iff it sorts after the current left head in (end,start,position)
                if (!has_left) dup = true;
                else dup = (r.e != l.e) ? r.e > l.e : (r.s != l.s) ? r.s > l.s : r.pos > l.pos;
            }
        }
        if (!include_overlaps && cand_distance(qs, qe, next.s, next.e) == 0) continue;  // nearest_index.rs:182-184
        if (!dup) emit(next);
    }
    return n_out;
}

template <bool EMIT>
__global__ void __launch_bounds__(PB_THREADS) nearest_k_kernel(IndexView ix, const uint32_t* __restrict__ key,
                                                               const int32_t* __restrict__ start,
                                                               const int32_t* __restrict__ end, uint64_t n, uint32_t k,
                                                               int include_overlaps, int strict,
                                                               uint32_t* __restrict__ counts,
                                                               const uint64_t* __restrict__ offsets,
                                                               uint32_t* __restrict__ build_idx,
                                                               uint32_t* __restrict__ valid,
                                                               uint32_t* __restrict__ probe_idx,
                                                               int64_t* __restrict__ dist) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t rs = ld_stream(start + i), re = ld_stream(end + i);
        uint32_t c = load_key(key, i);
        int32_t qs = rs, qe = re;
        if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); }
        uint64_t base = EMIT ? offsets[i] : 0;
        uint32_t m = 0;
        if (c < ix.n_keys)
            m = nearest_k_walk<EMIT>(ix, c, qs, qe, rs, re, k, include_overlaps, (uint32_t)i, base, build_idx, valid,
                                     probe_idx, dist);
        if (EMIT) {
            if (m == 0) {  // nearest.rs:433-439: one NULL row
                build_idx[base] = 0u;
                valid[base] = 0u;
                probe_idx[base] = (uint32_t)i;
                if (dist) dist[base] = 0;
            }
        } else {
            counts[i] = m ? m : 1u;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// subtract (subtract.rs:395-440): the pieces of a left interval [ls, le] that no right interval covers.  The reference
// sweeps the right intervals of the contig in (start, end) order with a cursor; gaps open exactly between consecutive
// MERGED right intervals (`next.start <= cur.end` merges, interval_tree.rs:52-73 -- the same rule as `rs > cursor`), so the
// sweep runs over the merged intervals of the coverage index:
//   skip runs ending before ls (strict: at or before), stop at runs starting after le (strict: at or after).
// ---------------------------------------------------------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(PB_THREADS) subtract_kernel(IndexView ix, const uint32_t* __restrict__ key,
                                                              const int32_t* __restrict__ start,
                                                              const int32_t* __restrict__ end, uint64_t n, int strict,
                                                              uint32_t* __restrict__ counts,
                                                              const uint64_t* __restrict__ offsets,
                                                              uint32_t* __restrict__ out_row, int32_t* __restrict__ out_start,
                                                              int32_t* __restrict__ out_end) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t ls = ld_stream(start + i), le = ld_stream(end + i);
        const uint32_t c = load_key(key, i);
        uint64_t o = EMIT ? offsets[i] : 0;
        uint32_t m = 0;
        int32_t cursor = ls;
        if (c < ix.n_keys) {
            const KeyDir kl = load_keydir(ix.d_mlast.kd, c), kf = load_keydir(ix.d_mfirst.kd, c);
            const uint32_t lo = strict ? rank_le_kd(ix.d_mlast, kl, ls) : rank_lt_kd(ix.d_mlast, kl, ls);
            const uint32_t hi = strict ? rank_lt_kd(ix.d_mfirst, kf, le) : rank_le_kd(ix.d_mfirst, kf, le);
            for (uint32_t j = lo; j < hi; j++) {
                const int4 r = __ldg(ix.m_rec + j);  // {first, last, ..}
                if (r.x > cursor) {
                    if (EMIT) { out_row[o] = (uint32_t)i; out_start[o] = cursor; out_end[o] = r.x; o++; }
                    m++;
                }
                if (r.y > cursor) cursor = r.y;
            }
        }
        if (cursor < le) {
            if (EMIT) { out_row[o] = (uint32_t)i; out_start[o] = cursor; out_end[o] = le; }
            m++;
        }
        if (!EMIT) counts[i] = m;
    }
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
// to-do list of deferred chunks: stream-ordered allocation, so concurrent calls on different streams never share it
static int todo_alloc(uint64_t rows, TodoList* td, cudaStream_t st) {
    const uint64_t chunks = (rows + TP_WTILE - 1) / TP_WTILE;
    void* p = nullptr;
    BRG_CUDA(cudaMallocAsync(&p, (chunks + 1) * sizeof(uint32_t), st));
    td->count = reinterpret_cast<uint32_t*>(p);
    td->chunks = td->count + 1;
    BRG_CUDA(cudaMemsetAsync(td->count, 0, sizeof(uint32_t), st));
    return 0;
}
static int todo_free(const TodoList& td, cudaStream_t st) {
    BRG_CUDA(cudaFreeAsync(td.count, st));
    return 0;
}
static inline unsigned todo_grid(uint64_t rows) {
    uint64_t chunks = (rows + TP_WTILE - 1) / TP_WTILE;
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(chunks, (uint64_t)kSMs * 8));
}
// The TMA-tiled kernels are the default for all three probes (measured on B200, bench extras: count 0.64 ms tiled vs
// 1.07 ms per-row; coverage 0.68 vs 0.69; nearest 0.72 vs 0.79).  BRG_TILED=<list> restricts the tiled path to the
// named probes (e.g. BRG_TILED=none runs everything on the per-row kernels) -- a debugging switch.
static bool tiled_enabled(const char* what, bool dflt) {
    static const char* env = getenv("BRG_TILED");
    if (!env) return dflt;
    return strstr(env, what) != nullptr;
}
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline unsigned probe_grid(uint64_t n, int per_sm) {
    // persistent-style grid: a multiple of the SM count, grid-stride over rows
    uint64_t want = (n + PB_THREADS - 1) / PB_THREADS;
    uint64_t cap = (uint64_t)kSMs * per_sm;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

// the windowed kernel for whole 256-row chunks (BRG_COUNT_WIN / BRG_COV_WIN select it), its to-do list served by `todo_kernel`
template <int OP>
static int launch_rank2_win(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t rows, int strict,
                            int64_t* out, const TodoList& todo, cudaStream_t st) {
    const int smem = (int)sizeof(Rank2Warp) * R2_WARPS;
    BRG_CUDA(cudaFuncSetAttribute((const void*)rank2_kernel_win<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const uint64_t chunks = rows / TP_WTILE;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((chunks + R2_WARPS - 1) / R2_WARPS, (uint64_t)kSMs * BRG_R2_CTAS));
    rank2_kernel_win<OP><<<grid, R2_WARPS * 32, smem, st>>>(ix, key, s, e, rows, strict, out, todo);
    BRG_LAUNCH_CHECK();
    return 0;
}
static bool win_enabled(const char* env_name, bool dflt) {
    const char* v = getenv(env_name);
    return v ? v[0] != '0' : dflt;
}

int launch_count(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n, int strict,
                 int64_t* out, cudaStream_t st) {
    if (n == 0) return 0;
    uint64_t done = 0;
    static const bool count_win = win_enabled("BRG_COUNT_WIN", BRG_COUNT_WIN_DEFAULT);
    if (count_win && tiled_enabled("count", true) && n >= TP_WTILE && ix.n_keys) {
        done = (n / TP_WTILE) * TP_WTILE;
        int rc = 0;
        TodoList todo;
        if ((rc = todo_alloc(done, &todo, st))) return rc;
        if ((rc = launch_rank2_win<0>(ix, key, s, e, done, strict, out, todo, st))) return rc;
        count_kernel_todo<<<todo_grid(done), TP_WTILE, 0, st>>>(ix, key, s, e, done, strict, out, todo);
        BRG_LAUNCH_CHECK();
        if ((rc = todo_free(todo, st))) return rc;
    } else if (tiled_enabled("count", true) && aligned16(key) && aligned16(s) && aligned16(e) && aligned16(out) && n >= PV) {
        done = (n / PV) * PV;
        static const int ct = [] { const char* v = getenv("BRG_COUNT_THREADS"); return v ? atoi(v) : BRG_COUNT_THREADS; }();
        int rc = 0;
        TodoList todo;
        if ((rc = todo_alloc(done, &todo, st))) return rc;
#define BRG_COUNT_LAUNCH(T)                                                                                                        \
    {                                                                                                                              \
        const int smem = (int)sizeof(WarpRing) * ((T) / 32);                                                                       \
        BRG_CUDA(cudaFuncSetAttribute((const void*)count_kernel_tiled<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));     \
        const uint64_t chunks = (done + TP_WTILE - 1) / TP_WTILE, per_cta = (T) / 32;                                              \
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((chunks + per_cta - 1) / per_cta, (uint64_t)kSMs * BRG_CC_CTAS)); \
        count_kernel_tiled<T><<<grid, (T), smem, st>>>(ix, key, s, e, done, strict, out, todo);                                    \
    }
        if (ct == 256) BRG_COUNT_LAUNCH(256)
        else BRG_COUNT_LAUNCH(320)
#undef BRG_COUNT_LAUNCH
        BRG_LAUNCH_CHECK();
        count_kernel_todo<<<todo_grid(done), TP_WTILE, 0, st>>>(ix, key, s, e, done, strict, out, todo);
        BRG_LAUNCH_CHECK();
        if ((rc = todo_free(todo, st))) return rc;
    }
    if (done < n) {  // unaligned buffers or the last n % PV rows
        count_kernel<<<probe_grid(n - done, 32), PB_THREADS, 0, st>>>(ix, key ? key + done : nullptr, s + done, e + done,
                                                                      n - done, strict, out + done);
        BRG_LAUNCH_CHECK();
    }
    return 0;
}
int launch_coverage(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                    int strict, int64_t* out, cudaStream_t st) {
    if (n == 0) return 0;
    uint64_t done = 0;
    static const bool cov_win = win_enabled("BRG_COV_WIN", BRG_COV_WIN_DEFAULT);
    if (cov_win && tiled_enabled("coverage", true) && n >= TP_WTILE && ix.n_keys) {
        done = (n / TP_WTILE) * TP_WTILE;
        int rc = 0;
        TodoList todo;
        if ((rc = todo_alloc(done, &todo, st))) return rc;
        if ((rc = launch_rank2_win<1>(ix, key, s, e, done, strict, out, todo, st))) return rc;
        coverage_kernel_todo<<<todo_grid(done), TP_WTILE, 0, st>>>(ix, key, s, e, done, strict, out, todo);
        BRG_LAUNCH_CHECK();
        if ((rc = todo_free(todo, st))) return rc;
    } else if (tiled_enabled("coverage", true) && aligned16(key) && aligned16(s) && aligned16(e) && aligned16(out) && n >= PV) {
        done = (n / PV) * PV;
        static const int ct = [] { const char* v = getenv("BRG_COVERAGE_THREADS"); return v ? atoi(v) : BRG_COVERAGE_THREADS; }();
        int rc = 0;
        TodoList todo;
        if ((rc = todo_alloc(done, &todo, st))) return rc;
#define BRG_COV_LAUNCH(T)                                                                                                          \
    {                                                                                                                              \
        const int smem = (int)sizeof(WarpRing) * ((T) / 32);                                                                       \
        BRG_CUDA(cudaFuncSetAttribute((const void*)coverage_kernel_tiled<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));  \
        const uint64_t chunks = (done + TP_WTILE - 1) / TP_WTILE, per_cta = (T) / 32;                                              \
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((chunks + per_cta - 1) / per_cta, (uint64_t)kSMs * BRG_CC_CTAS)); \
        coverage_kernel_tiled<T><<<grid, (T), smem, st>>>(ix, key, s, e, done, strict, out, todo);                                 \
    }
        if (ct == 256) BRG_COV_LAUNCH(256)
        else BRG_COV_LAUNCH(320)
#undef BRG_COV_LAUNCH
        BRG_LAUNCH_CHECK();
        coverage_kernel_todo<<<todo_grid(done), TP_WTILE, 0, st>>>(ix, key, s, e, done, strict, out, todo);
        BRG_LAUNCH_CHECK();
        if ((rc = todo_free(todo, st))) return rc;
    }
    if (done < n) {
        coverage_kernel<<<probe_grid(n - done, 32), PB_THREADS, 0, st>>>(ix, key ? key + done : nullptr, s + done, e + done,
                                                                         n - done, strict, out + done);
        BRG_LAUNCH_CHECK();
    }
    return 0;
}
int launch_nearest_k1(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                      int include_overlaps, int strict, uint32_t* left_idx, uint32_t* valid_bits, int64_t* dist,
                      cudaStream_t st) {
    if (n == 0) return 0;
    uint64_t done = 0;
    if (tiled_enabled("nearest", true) && aligned16(key) && aligned16(s) && aligned16(e) && aligned16(left_idx) && aligned16(dist) &&
        n >= TP_WTILE) {
        done = (n / TP_WTILE) * TP_WTILE;  // whole chunks only: validity words never straddle the two kernels
        static const int ct = [] { const char* v = getenv("BRG_NEAREST_THREADS"); return v ? atoi(v) : BRG_NEAREST_THREADS; }();
        int rc = 0;
        TodoList todo;
        if ((rc = todo_alloc(done, &todo, st))) return rc;
#define BRG_NEAR_LAUNCH(T)                                                                                                         \
    {                                                                                                                              \
        const int smem = (int)sizeof(WarpRing) * ((T) / 32);                                                                       \
        BRG_CUDA(cudaFuncSetAttribute((const void*)nearest_k1_kernel_tiled<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        const uint64_t chunks = (done + TP_WTILE - 1) / TP_WTILE, per_cta = (T) / 32;                                              \
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((chunks + per_cta - 1) / per_cta, (uint64_t)kSMs * 2)); \
        nearest_k1_kernel_tiled<T><<<grid, (T), smem, st>>>(ix, key, s, e, done, include_overlaps, strict, left_idx,               \
                                                            reinterpret_cast<uint8_t*>(valid_bits), dist, todo);                   \
    }
        static const bool windowed = [] { const char* v = getenv("BRG_NEAREST_WIN"); return !v || v[0] != '0'; }();
        if (windowed) {
            const int smem = (int)sizeof(NearWarp) * NW_WARPS;
            BRG_CUDA(cudaFuncSetAttribute((const void*)nearest_k1_kernel_win, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            const uint64_t chunks = done / TP_WTILE;
            const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((chunks + NW_WARPS - 1) / NW_WARPS, (uint64_t)kSMs * BRG_NW_CTAS));
            nearest_k1_kernel_win<<<grid, NW_WARPS * 32, smem, st>>>(ix, key, s, e, done, include_overlaps, strict, left_idx, valid_bits, dist, todo);
        } else if (ct == 320) BRG_NEAR_LAUNCH(320)  // (spills: 0.87 ms against 0.57 ms; kept as the measured alternative)
        else BRG_NEAR_LAUNCH(256)
#undef BRG_NEAR_LAUNCH
        BRG_LAUNCH_CHECK();
        nearest_k1_kernel_todo<<<todo_grid(done), TP_WTILE, 0, st>>>(ix, key, s, e, done, include_overlaps, strict, left_idx,
                                                                     valid_bits, dist, todo);
        BRG_LAUNCH_CHECK();
        if ((rc = todo_free(todo, st))) return rc;
    }
    if (done < n) {
        nearest_k1_kernel<<<probe_grid(n - done, 32), PB_THREADS, 0, st>>>(
            ix, key ? key + done : nullptr, s + done, e + done, n - done, include_overlaps, strict, left_idx + done,
            valid_bits + done / 32, dist ? dist + done : nullptr);
        BRG_LAUNCH_CHECK();
    }
    return 0;
}
int launch_subtract(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n, int strict,
                    uint32_t* counts, const uint64_t* offsets, uint32_t* out_row, int32_t* out_start, int32_t* out_end,
                    cudaStream_t st) {
    if (n == 0) return 0;
    if (offsets)
        subtract_kernel<true><<<probe_grid(n, 32), PB_THREADS, 0, st>>>(ix, key, s, e, n, strict, nullptr, offsets, out_row, out_start, out_end);
    else
        subtract_kernel<false><<<probe_grid(n, 32), PB_THREADS, 0, st>>>(ix, key, s, e, n, strict, counts, nullptr, nullptr, nullptr, nullptr);
    BRG_LAUNCH_CHECK();
    return 0;
}
int launch_nearest_k_count(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                           uint32_t k, int include_overlaps, int strict, uint32_t* counts, cudaStream_t st) {
    if (n == 0) return 0;
    nearest_k_kernel<false><<<probe_grid(n, 32), PB_THREADS, 0, st>>>(ix, key, s, e, n, k, include_overlaps, strict,
                                                                     counts, nullptr, nullptr, nullptr, nullptr, nullptr);
    BRG_LAUNCH_CHECK();
    return 0;
}
int launch_nearest_k_emit(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                          uint32_t k, int include_overlaps, int strict, const uint64_t* offsets, uint32_t* build_idx,
                          uint32_t* valid, uint32_t* probe_idx, int64_t* dist, cudaStream_t st) {
    if (n == 0) return 0;
    nearest_k_kernel<true><<<probe_grid(n, 32), PB_THREADS, 0, st>>>(ix, key, s, e, n, k, include_overlaps, strict,
                                                                    nullptr, offsets, build_idx, valid, probe_idx, dist);
    BRG_LAUNCH_CHECK();
    return 0;
}

}  // namespace brg
