// brg_tile.cuh -- tile pipeline and warp-window rank search shared by the tiled probe kernels (sm_100a).
//
// Data movement: the probe columns are streamed with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx,
// L2 evict-first) into a per-warp shared-memory ring, issued by one lane, so HBM requests stay in flight independently
// of what the warp is doing.  (The first version of the kernels issued per-thread 4-byte loads and was latency-bound:
// 29 % DRAM throughput with every warp stalled on its probe loads, profiles/r1_count_v0_ncu.txt.)
//
// Search: every thread owns PV consecutive probe rows, a warp owns a contiguous range of 256-row chunks.  Genomic probe
// tables are coordinate-sorted (the reference relies on that locality too, nearest.rs:320-322), so the ranks of one
// chunk fall into a short window of the sorted build array and the window of the next chunk starts where the previous
// one ended.  Per sorted array the warp keeps a cursor (WarpCursor).  A chunk then costs
//   * one coalesced load of the next few dozen build values (they ARE the window) + a ballot to find its end,
//   * an INVERSE search: each of the (few) window elements binary-searches the (many) probe values of the chunk in
//     shared memory, drops a 1 into a 256-bin shared histogram at its insertion point, and an inclusive prefix sum
//     of the histogram (8 serial adds per lane + one warp scan) yields the rank of all 256 rows at once.
// That is ~100 instructions per lane for 8 rows instead of 8 independent searches.  It needs the chunk's query values
// to be non-decreasing (checked with one vote).  Anything that breaks the pattern -- unsorted rows, a key change inside
// the chunk, a partial chunk, a window wider than TP_WMAX (a build side about as dense as the probe side) -- makes the
// warp defer the chunk (or, for a wide window, its whole remaining range) to a one-row-per-thread kernel through a
// to-do list, so results are exact for any input.
//
// Shared-memory layout: these kernels are bound by the shared-memory pipe, so everything a lane touches with 128-bit
// accesses or that a binary search probes (query columns, histograms, output staging) lives at word i + 4 * (i / 32)
// (hist_slot / pad_tile / PadStep below): conflict-free for both access patterns.  Histograms have one byte per bin.
// Unsorted ends (variable-length probes) are ranked by a forward search over the staged window (warp_forward_ranks).
#pragma once
#include "brg_internal.cuh"

namespace brg {

constexpr int TP_THREADS = 256;
constexpr int TP_WARPS = TP_THREADS / 32;
constexpr int PV = 8;                       // consecutive probe rows per thread
constexpr int TP_WTILE = 32 * PV;           // 256 consecutive rows per warp chunk
constexpr int TP_TILE = TP_THREADS * PV;    // rows in flight per CTA step
constexpr int TP_STAGES = 2;
constexpr int TP_WCAP = 288;                // words of per-array scratch (256 histogram bins, padded / window copy of the fallback)
#ifndef BRG_NWIN
#define BRG_NWIN 3
#endif
constexpr int TP_NWIN = BRG_NWIN;                  // scratch buffers per warp (one per sorted array searched)
constexpr int TP_WMAX = 128;                // widest window (elements) the inverse search walks; beyond it the chunk is
                                            // deferred: with the build side about as dense as the probe side the
                                            // one-row-per-thread kernels win (10M x 10M sorted, B200: nearest 0.16 ms vs
                                            // 0.34 ms for a staged-window search, coverage 0.083 vs 0.113, count 0.16 vs 0.16)
constexpr uint32_t TP_BAIL = 3;             // consecutive deferred chunks after which a warp stops trying windows

#ifdef __CUDACC__
// ---- TMA / mbarrier primitives (inline PTX) -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk copy global -> shared, completion signalled on `bar` (bytes must be a multiple of 16, both 16 B aligned)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// Chunks the tiled kernels cannot serve from a window (mixed keys, unsorted rows) are appended to a to-do list and
// answered afterwards by a one-row-per-thread kernel that searches the rank directory at full occupancy.
struct TodoList {
    uint32_t* count;
    uint32_t* chunks;
};
__device__ __forceinline__ void todo_push(const TodoList& td, uint64_t chunk) {
    if ((threadIdx.x & 31) == 0) td.chunks[atomicAdd(td.count, 1u)] = (uint32_t)chunk;
}

// ---- shared-memory layout ----------------------------------------------------------------------------------------
// Every warp runs its own pipeline over its own contiguous range of 256-row chunks: a private 2-stage ring fed by
// TMA bulk copies that lane 0 issues, a private mbarrier per stage, private window buffers.  Warps never wait for
// each other (no __syncthreads in the main loop).
struct __align__(128) WarpRing {
    int32_t start[TP_STAGES][TP_WCAP];  // 256 rows land here linearly (TMA); pad_tile() rewrites them in the padded layout
    int32_t end[TP_STAGES][TP_WCAP];
    uint32_t key[TP_STAGES][TP_WTILE];
    uint32_t aux[TP_NWIN][TP_WCAP];  // per array: histogram of the inverse search / window copy of the fallback
    uint64_t full[TP_STAGES];
    uint64_t pad_[14];
};
struct __align__(128) ProbeRing {
    WarpRing w[TP_WARPS];
};

struct ProbeTile {
    int32_t qs[PV], qe[PV];
    uint32_t key_and, key_or;  // AND / OR over some of the chunk's keys (all of them once reduced over the warp)
    uint64_t row0;   // global row of this thread's first probe
    uint64_t chunk;  // index of this 256-row chunk
    bool active;     // thread owns rows in this chunk
    bool full;       // all 32 lanes own rows (256-row chunk)
    int32_t* s_qs;   // the chunk's start / end columns in shared memory (valid until release())
    int32_t* s_qe;
};

// Usage:  ProbePipe pipe(...); while (pipe.next(tile)) { ...the warp stays converged... }
struct ProbePipe {
    WarpRing* ring;
    const uint32_t* key;
    const int32_t* start;
    const int32_t* end;
    uint64_t n_rows;       // multiple of PV
    uint64_t chunk_begin;  // this warp's 256-row chunks: [chunk_begin, chunk_end)
    uint64_t chunk_end;
    uint64_t it;
    uint64_t policy;

    __device__ __forceinline__ void issue(uint64_t iter) {
        const uint64_t c = chunk_begin + iter;
        if (c >= chunk_end) return;
        const int s = (int)(iter % TP_STAGES);
        const uint64_t row = c * TP_WTILE;
        const uint32_t rows = (uint32_t)min((uint64_t)TP_WTILE, n_rows - row);
        const uint32_t bytes = rows * 4u;
        mbar_expect_tx(&ring->full[s], bytes * (key ? 3u : 2u));
        tma_load_1d(ring->start[s], start + row, bytes, &ring->full[s], policy);
        tma_load_1d(ring->end[s], end + row, bytes, &ring->full[s], policy);
        if (key) tma_load_1d(ring->key[s], key + row, bytes, &ring->full[s], policy);
    }

    __device__ __forceinline__ ProbePipe(ProbeRing* r, const uint32_t* k, const int32_t* s, const int32_t* e, uint64_t rows)
        : ProbePipe(r->w, k, s, e, rows) {}
    // rings: one WarpRing per warp of the CTA (any CTA size)
    __device__ __forceinline__ ProbePipe(WarpRing* rings, const uint32_t* k, const int32_t* s, const int32_t* e,
                                         uint64_t rows)
        : ring(&rings[threadIdx.x >> 5]), key(k), start(s), end(e), n_rows(rows), it(0) {
        const uint64_t n_chunks = (rows + TP_WTILE - 1) / TP_WTILE;
        const uint64_t cta_warps = blockDim.x >> 5;
        const uint64_t n_warps = (uint64_t)gridDim.x * cta_warps;
        const uint64_t wid = (uint64_t)blockIdx.x * cta_warps + (threadIdx.x >> 5);
        const uint64_t q = n_chunks / n_warps, rem = n_chunks % n_warps;
        chunk_begin = wid * q + min(wid, rem);
        chunk_end = chunk_begin + q + (wid < rem ? 1 : 0);
        const int lane = threadIdx.x & 31;
        if (lane == 0) {
            for (int i = 0; i < TP_STAGES; i++) mbar_init(&ring->full[i], 1);
            mbar_fence_init();
        }
        __syncwarp();
        if (lane == 0) {
            policy = l2_evict_first_policy();
            for (int i = 0; i < TP_STAGES; i++) issue(i);
        }
    }

    __device__ __forceinline__ bool next(ProbeTile& tile) {
        const uint64_t c = chunk_begin + it;
        if (c >= chunk_end) return false;
        const int s = (int)(it % TP_STAGES);
        mbar_wait(&ring->full[s], (uint32_t)((it / TP_STAGES) & 1));
        const uint64_t row = c * TP_WTILE;
        const uint32_t rows = (uint32_t)min((uint64_t)TP_WTILE, n_rows - row);
        const uint32_t my = (threadIdx.x & 31) * PV;
        tile.active = my < rows;
        tile.row0 = row + my;
        tile.chunk = c;
        if (tile.active) {
            const int4* ps = reinterpret_cast<const int4*>(&ring->start[s][my]);
            const int4* pe = reinterpret_cast<const int4*>(&ring->end[s][my]);
            int4 a = ps[0], b = ps[1], c4 = pe[0], d = pe[1];
            tile.qs[0] = a.x; tile.qs[1] = a.y; tile.qs[2] = a.z; tile.qs[3] = a.w;
            tile.qs[4] = b.x; tile.qs[5] = b.y; tile.qs[6] = b.z; tile.qs[7] = b.w;
            tile.qe[0] = c4.x; tile.qe[1] = c4.y; tile.qe[2] = c4.z; tile.qe[3] = c4.w;
            tile.qe[4] = d.x; tile.qe[5] = d.y; tile.qe[6] = d.z; tile.qe[7] = d.w;
        } else {
#pragma unroll
            for (int j = 0; j < PV; j++) { tile.qs[j] = 0; tile.qe[j] = 0; }
        }
        // The keys are only checked for uniformity, so which lane looks at which rows does not matter: lane L reads the 16-byte
        // units L and L + 32 (conflict-free; the 8 consecutive rows it owns would cost two wavefronts per load).
        uint32_t ka = 0xFFFFFFFFu, ko = 0u;
        if (key) {
            const uint4* pk = reinterpret_cast<const uint4*>(ring->key[s]);
            const uint32_t lane = threadIdx.x & 31;
            if (4u * lane < rows) { const uint4 k0 = pk[lane]; ka &= k0.x & k0.y & k0.z & k0.w; ko |= k0.x | k0.y | k0.z | k0.w; }
            if (128u + 4u * lane < rows) { const uint4 k1 = pk[32 + lane]; ka &= k1.x & k1.y & k1.z & k1.w; ko |= k1.x | k1.y | k1.z | k1.w; }
        } else if (tile.active) {
            ka = 0u;  // no key column: every row has key 0
        }
        tile.key_and = ka;
        tile.key_or = ko;
        tile.full = rows == (uint32_t)TP_WTILE;
        tile.s_qs = ring->start[s];
        tile.s_qe = ring->end[s];
        return true;
    }
    // the chunk's shared-memory columns are no longer needed: refill the stage with the chunk after next
    __device__ __forceinline__ void release() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) issue(it + TP_STAGES);
        it++;
    }
    // Give up on the windowed search (the input is evidently not coordinate-sorted): hand every remaining chunk of this
    // warp to the per-row kernel.  Call instead of release().  The bulk copies already in flight must land before the
    // CTA may exit, so their barriers are still waited for.
    __device__ __forceinline__ void bail_out(const TodoList& td) {
        __syncwarp();
        for (uint64_t i = 1; i < (uint64_t)TP_STAGES && chunk_begin + it + i < chunk_end; i++)
            mbar_wait(&ring->full[(it + i) % TP_STAGES], (uint32_t)(((it + i) / TP_STAGES) & 1));
        const uint64_t first = chunk_begin + it + 1;
        if (first < chunk_end) {
            const uint32_t n_left = (uint32_t)(chunk_end - first);
            const int lane = threadIdx.x & 31;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(td.count, n_left);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            for (uint32_t i = lane; i < n_left; i += 32) td.chunks[base + i] = (uint32_t)(first + i);
        }
        it = chunk_end - chunk_begin;
    }
};

// strict filter: probe start+1 / end-1 (interval_tree.rs:253-256, nearest.rs:341-344); registers only -- pad_tile() below
// brings the shared copy of the columns, which the inverse search reads, in step
__device__ __forceinline__ void strict_adjust(ProbeTile& t) {
#pragma unroll
    for (int j = 0; j < PV; j++) { t.qs[j] = wadd(t.qs[j], 1); t.qe[j] = wsub(t.qe[j], 1); }
}

// All keys of the chunk equal one valid key?  (AND == OR over all of them; rows past the end contribute neutral values.)
__device__ __forceinline__ bool warp_uniform_key(const ProbeTile& t, uint32_t n_keys, uint32_t* kout) {
    const uint32_t a = __reduce_and_sync(0xFFFFFFFFu, t.key_and);
    const uint32_t o = __reduce_or_sync(0xFFFFFFFFu, t.key_or);
    *kout = o;
    return a == o && o < n_keys;  // (no rows: a = ~0, o = 0 -> false)
}

// Histogram bins live at word i + 4 * (i / 32): 16 pad bytes after every 128.  A lane zeroes and reads back its 8 bins with
// two 128-bit accesses; unpadded, lanes L and L + 4 of a quarter warp hit the same banks (every such access takes two
// wavefronts) -- and the tiled kernels are bound by the shared-memory pipe (ncu: l1tex data-pipe wavefronts at 86 % of peak,
// profiles/r1_count_v14_ncu.txt), not by DRAM or issue slots.
__device__ __forceinline__ uint32_t hist_slot(uint32_t bin) { return bin + ((bin >> 5) << 2); }
__device__ __forceinline__ uint4* hist_lane_bins(uint32_t* aux, int lane) { return reinterpret_cast<uint4*>(aux) + 2 * lane + (lane >> 2); }

// The histogram of an inverse search has one BYTE per padded query slot (288 bytes): a window holds at most TP_WMAX = 128
// elements, so no bin overflows and the packed prefix sums below stay below 256.  Against 32-bit bins: zeroing is one
// 128-bit store by 18 lanes (3 wavefronts instead of 8), a lane reads its 8 bins with two 32-bit loads (4 instead of 8), and
// the serial part of the prefix sum is two multiplications.
__device__ __forceinline__ void hist8_zero(uint32_t* aux, int lane) {
    if (lane < 18) reinterpret_cast<uint4*>(aux)[lane] = make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void hist8_add(uint32_t* aux, uint32_t padded_slot) {
    atomicAdd(&aux[padded_slot >> 2], 1u << ((padded_slot & 3u) << 3));
}
// inclusive prefix sums (within the lane) of the lane's 8 bins, and their total
__device__ __forceinline__ void hist8_lane_prefix(const uint32_t* aux, int lane, uint32_t (&c)[PV], uint32_t* total) {
    const uint32_t w = 2 * lane + (lane >> 2);  // the lane's 8 bins start at padded slot 8 * lane + 4 * (lane / 4)
    const uint32_t p0 = aux[w] * 0x01010101u;
    const uint32_t p1 = (aux[w + 1] + (p0 >> 24)) * 0x01010101u;  // (the first half's total rides in byte 0: it enters every prefix)
    c[0] = p0 & 0xFFu; c[1] = (p0 >> 8) & 0xFFu; c[2] = (p0 >> 16) & 0xFFu; c[3] = p0 >> 24;
    c[4] = p1 & 0xFFu; c[5] = (p1 >> 8) & 0xFFu; c[6] = (p1 >> 16) & 0xFFu; c[7] = p1 >> 24;
    *total = c[7];
}

// Rewrites the chunk's start / end columns in shared memory (in place, from the registers) in the padded layout of hist_slot:
// word i at i + 4 * (i / 32).  The binary searches of the inverse search probe Q[pos + step - 1] with pos a multiple of
// 2 * step: on a linear layout the probes of one step fall into 1, 1, 1, 1, 2, 4, 8, 16 banks (up to 8 wavefronts per
// load); padded, a step's probes spread over 8 to 32 banks.  Two conflict-free 128-bit stores per column and lane.
// Call once per chunk after next() (and after strict_adjust); Q below always means this layout.
__device__ __forceinline__ void pad_tile(ProbeTile& t) {
    __syncwarp();  // every lane holds its rows in registers
    if (t.active) {
        const int lane = threadIdx.x & 31;
        uint4* ps = hist_lane_bins(reinterpret_cast<uint32_t*>(t.s_qs), lane);
        uint4* pe = hist_lane_bins(reinterpret_cast<uint32_t*>(t.s_qe), lane);
        ps[0] = make_uint4(t.qs[0], t.qs[1], t.qs[2], t.qs[3]); ps[1] = make_uint4(t.qs[4], t.qs[5], t.qs[6], t.qs[7]);
        pe[0] = make_uint4(t.qe[0], t.qe[1], t.qe[2], t.qe[3]); pe[1] = make_uint4(t.qe[4], t.qe[5], t.qe[6], t.qe[7]);
    }
    // (These stores go to a buffer that the TMA refill issued by release() overwrites later.  No proxy fence in between: every
    //  lane's stores are read back by the warp's searches -- hence performed -- long before release() runs, and the refill is
    //  a write, which observes nothing.  A fence.proxy.async here was measured at +1.7 % kernel time.)
    __syncwarp();
}
// One step of a binary search over the 256 padded queries, in padded coordinates: `pos` is the padded slot of a multiple of
// 2 * STEP, the probe is linear element pos + STEP - 1, the advance is STEP elements.
template <uint32_t STEP> struct PadStep {
    static constexpr uint32_t probe = STEP >= 32 ? STEP + 4 * (STEP / 32) - 5 : STEP - 1;  // (the element before a block boundary)
    static constexpr uint32_t advance = STEP >= 32 ? STEP + 4 * (STEP / 32) : STEP;
};
// padded insertion slot of v among the chunk's queries:  LT = false: #(Q < v),  LT = true: #(Q <= v)   (0 .. 255, padded)
template <bool LT>
__device__ __forceinline__ uint32_t padded_insertion_slot(const int32_t* Q, int32_t v) {
    uint32_t pos = 0;
#define BRG_PAD_STEP(S) { const int32_t qv = Q[pos + PadStep<S>::probe]; pos += (LT ? (qv <= v) : (qv < v)) ? PadStep<S>::advance : 0u; }
    BRG_PAD_STEP(128) BRG_PAD_STEP(64) BRG_PAD_STEP(32) BRG_PAD_STEP(16) BRG_PAD_STEP(8) BRG_PAD_STEP(4) BRG_PAD_STEP(2) BRG_PAD_STEP(1)
#undef BRG_PAD_STEP
    return pos;
}

// ---- warp cursor over one sorted array ---------------------------------------------------------------------------
// Invariant while valid: every element of key `key`'s segment before `pos` is <= bound (LT: < bound), hence `pos` is a
// legal window start for any tile whose smallest query is >= bound.
template <bool FAR>
struct WarpCursorT {
    static constexpr bool kFar = FAR;
    uint32_t pos;
    int32_t bound;
    uint32_t key;
    bool valid;
    bool wide;    // the last failure was a window wider than TP_WMAX: a property of the region, not of one chunk
    int32_t pre;  // prefetched A[pos + lane] (or +inf past the segment): the first 32 window candidates of the next chunk
    int32_t far;  // FAR only: prefetched A[pos + TP_WMAX - 1] (or +inf): decides "window too wide" without a dependent load
    __device__ __forceinline__ void reset() { valid = false; wide = false; pos = 0; bound = 0; key = 0xFFFFFFFFu; pre = INT32_MAX; far = INT32_MAX; }
    // (after pos changed) loads what the next chunk's search looks at first
    __device__ __forceinline__ void prefetch(const int32_t* A, uint32_t seg_end, int lane) {
        pre = (pos + lane < seg_end) ? __ldg(A + pos + lane) : INT32_MAX;
        if (FAR) far = (pos + (uint32_t)TP_WMAX - 1 < seg_end) ? __ldg(A + pos + (uint32_t)TP_WMAX - 1) : INT32_MAX;
    }
};
// The far prefetch pays where cursors are served by the single search (nearest: 0.767 -> 0.700 ms); kernels whose cursors
// go through the interleaved dual search, which never looks at the far element, lose 1-2 % (and registers) to it.
using WarpCursor = WarpCursorT<false>;
using WarpCursorFar = WarpCursorT<true>;

// are the chunk's 256 query values non-decreasing in row order?  (only full chunks qualify)
__device__ __forceinline__ bool chunk_mono(const int32_t (&q)[PV], bool full) {
    if (!full) return false;
    bool up = true;
#pragma unroll
    for (int j = 1; j < PV; j++) up &= q[j] >= q[j - 1];
    const int32_t prev_last = __shfl_up_sync(0xFFFFFFFFu, q[PV - 1], 1);
    up &= ((threadIdx.x & 31) == 0) | (q[0] >= prev_last);
    return __all_sync(0xFFFFFFFFu, up);
}

// r[j] = global index of the first element of key k's segment that is  > q[j] (LT = false)  or  >= q[j] (LT = true).
// Must be called by all 32 lanes; `k`/`kid` warp-uniform.  Lanes with active == false contribute nothing and get
// unspecified ranks.  Returns false when the chunk's window does not fit the shared buffers (the caller defers the chunk
// to the per-row kernel).  mono = chunk_mono(q).  Q = the chunk's 256 query values in shared memory (same values as q,
// row-major), only read when mono.  aux = this warp's 256-word scratch for this array.
template <bool LT, typename CUR>
__device__ __forceinline__ bool warp_window_ranks(CUR& cur, const DirView& d, const KeyDir& k, uint32_t kid,
                                                  const int32_t (&q)[PV], bool active, bool mono, const int32_t* Q,
                                                  uint32_t* aux, uint32_t (&r)[PV]) {
    const int lane = threadIdx.x & 31;
    int32_t mn, mx;
    if (mono) {
        mn = __shfl_sync(0xFFFFFFFFu, q[0], 0);
        mx = __shfl_sync(0xFFFFFFFFu, q[PV - 1], 31);
    } else {
        mn = INT32_MAX; mx = INT32_MIN;
        if (active) {
#pragma unroll
            for (int j = 0; j < PV; j++) { mn = min(mn, q[j]); mx = max(mx, q[j]); }
        }
        mn = __reduce_min_sync(0xFFFFFFFFu, mn);
        mx = __reduce_max_sync(0xFFFFFFFFu, mx);
        if (mn > mx) return true;  // no active lane in this warp
    }
    // ---- window start: the cursor when the queries moved forward, else a directory lookup ------------------------
    uint32_t lo;
    const bool reuse = cur.valid && cur.key == kid && mn >= cur.bound;
    if (reuse) lo = cur.pos;
    else lo = LT ? rank_lt_kd(d, k, mn) : rank_le_kd(d, k, mn);

    if (mono) {
        // ---- too wide for the inverse search?  One load decides, before any work is spent on the chunk: the element
        // TP_WMAX places after the window start.  (Dense build sides are the per-row kernels' job, see TP_WMAX.)
        {
            const uint32_t far = lo + (uint32_t)TP_WMAX - 1;
            if (far < k.seg_end) {
                const int32_t vf = (CUR::kFar && reuse) ? cur.far : __ldg(d.A + far);  // (reuse: loaded while the previous chunk was being finished)
                if (LT ? vf < mx : vf <= mx) { cur.valid = false; cur.wide = true; return false; }
            }
        }
        // ---- inverse search ---------------------------------------------------------------------------------------
        hist8_zero(aux, lane);
        __syncwarp();
        uint32_t wn = 0;
        bool open = true;
#pragma unroll 1
        for (uint32_t base = 0; open && base < (uint32_t)TP_WMAX; base += 32) {
            const uint32_t idx = lo + base + lane;
            int32_t v;
            if (reuse && base == 0) v = cur.pre;  // loaded while the previous chunk was being finished
            else v = idx < k.seg_end ? __ldg(d.A + idx) : INT32_MAX;
            const bool in = idx < k.seg_end && (LT ? v < mx : v <= mx);
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, in);
            wn += __popc(m);
            open = (m == 0xFFFFFFFFu);
            if (m) {
                // insertion point of v among the 256 sorted queries:  LE: #(Q < v)   LT: #(Q <= v)
                // (the 8 steps cover Q[0..254]; an in-window element never lies above Q[255] = mx, so pos <= 255 is exact)
                const uint32_t slot = padded_insertion_slot<LT>(Q, v);
                if (in) hist8_add(aux, slot);
            }
        }
        __syncwarp();
        if (!open) {
            cur.pos = lo + wn; cur.bound = mx; cur.key = kid; cur.valid = true;
            cur.prefetch(d.A, k.seg_end, lane);
            uint32_t c[PV], incl;
            hist8_lane_prefix(aux, lane, c, &incl);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += y;
            }
            const uint32_t before = lo + incl - c[PV - 1];
#pragma unroll
            for (int j = 0; j < PV; j++) r[j] = before + c[j];
            __syncwarp();
            return true;
        }
        __syncwarp();
    } else {
        // ---- per-row search over a shared copy of the window ------------------------------------------------------
        const uint32_t hi = LT ? rank_lt_kd(d, k, mx) : rank_le_kd(d, k, mx);
        const uint32_t wn = hi - lo;
        cur.pos = hi; cur.bound = mx; cur.key = kid; cur.valid = true;
        cur.prefetch(d.A, k.seg_end, lane);
        if (wn == 0) {
#pragma unroll
            for (int j = 0; j < PV; j++) r[j] = lo;
            return true;
        }
        if (wn < (uint32_t)TP_WCAP / 2) {
            int32_t* w = reinterpret_cast<int32_t*>(aux);
            const uint32_t top = 1u << (31 - __clz(wn));  // padded with +inf up to 2*top-1 entries: no bound checks
            for (uint32_t t = lane; t < 2 * top - 1; t += 32) w[t] = (t < wn) ? __ldg(d.A + lo + t) : INT32_MAX;
            __syncwarp();
            uint32_t pos[PV];
#pragma unroll
            for (int j = 0; j < PV; j++) pos[j] = 0;
            for (uint32_t step = top; step; step >>= 1) {
#pragma unroll
                for (int j = 0; j < PV; j++) {
                    const int32_t v = w[pos[j] + step - 1];
                    pos[j] += (LT ? (v < q[j]) : (v <= q[j])) ? step : 0u;
                }
            }
#pragma unroll
            for (int j = 0; j < PV; j++) r[j] = lo + min(pos[j], wn);  // q == INT32_MAX can count the padding
            __syncwarp();
            return true;
        }
    }
    // ---- window wider than the shared buffer (unsorted probes): the caller defers the chunk to the per-row kernel ---
    cur.valid = false;
    return false;
}

// Two inverse searches at once.  The common case of the sparse regime -- both windows fit one round of 32 elements --
// is a chain of dependent shared-memory accesses (8 search steps, histogram atomics, prefix sums): running the chains
// of the two arrays interleaved doubles the instruction-level parallelism of the warp.  Returns false WITHOUT touching
// cursors or results when the case does not apply (a cursor that cannot be reused, a window wider than one round); the caller
// then runs the two searches one after the other.  Preconditions: full chunk, q0 and q1 non-decreasing.
template <bool LT0, bool LT1, typename CUR0, typename CUR1>
__device__ __forceinline__ bool warp_window_ranks2(CUR0& c0, const DirView& d0, const KeyDir& k0, const int32_t (&q0)[PV],
                                                   const int32_t* Q0, uint32_t* aux0, uint32_t (&r0)[PV], CUR1& c1,
                                                   const DirView& d1, const KeyDir& k1, const int32_t (&q1)[PV],
                                                   const int32_t* Q1, uint32_t* aux1, uint32_t (&r1)[PV], uint32_t kid) {
    const int lane = threadIdx.x & 31;
    const int32_t mn0 = __shfl_sync(0xFFFFFFFFu, q0[0], 0), mx0 = __shfl_sync(0xFFFFFFFFu, q0[PV - 1], 31);
    const int32_t mn1 = __shfl_sync(0xFFFFFFFFu, q1[0], 0), mx1 = __shfl_sync(0xFFFFFFFFu, q1[PV - 1], 31);
    const bool reuse0 = c0.valid && c0.key == kid && mn0 >= c0.bound;
    const bool reuse1 = c1.valid && c1.key == kid && mn1 >= c1.bound;
    if (!(reuse0 && reuse1)) return false;  // (first chunk of a run: the single searches consult the directory)
    const uint32_t lo0 = c0.pos, lo1 = c1.pos;
    const int32_t v0 = c0.pre, v1 = c1.pre;
    const bool in0 = lo0 + lane < k0.seg_end && (LT0 ? v0 < mx0 : v0 <= mx0);
    const bool in1 = lo1 + lane < k1.seg_end && (LT1 ? v1 < mx1 : v1 <= mx1);
    const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
    if (m0 == 0xFFFFFFFFu || m1 == 0xFFFFFFFFu) return false;  // a window continues past this round
    hist8_zero(aux0, lane);
    hist8_zero(aux1, lane);
    __syncwarp();
    if (m0 | m1) {
        uint32_t p0 = 0, p1 = 0;  // padded slots
#define BRG_PAD_STEP2(S)                                                                              \
    {                                                                                                 \
        const int32_t x0 = Q0[p0 + PadStep<S>::probe], x1 = Q1[p1 + PadStep<S>::probe];               \
        p0 += (LT0 ? (x0 <= v0) : (x0 < v0)) ? PadStep<S>::advance : 0u;                              \
        p1 += (LT1 ? (x1 <= v1) : (x1 < v1)) ? PadStep<S>::advance : 0u;                              \
    }
        BRG_PAD_STEP2(128) BRG_PAD_STEP2(64) BRG_PAD_STEP2(32) BRG_PAD_STEP2(16) BRG_PAD_STEP2(8) BRG_PAD_STEP2(4) BRG_PAD_STEP2(2) BRG_PAD_STEP2(1)
#undef BRG_PAD_STEP2
        if (in0) hist8_add(aux0, p0);
        if (in1) hist8_add(aux1, p1);
    }
    __syncwarp();
    c0.pos = lo0 + __popc(m0); c0.bound = mx0;
    c1.pos = lo1 + __popc(m1); c1.bound = mx1;
    c0.prefetch(d0.A, k0.seg_end, lane);
    c1.prefetch(d1.A, k1.seg_end, lane);
    uint32_t e0[PV], e1[PV], i0, i1;
    hist8_lane_prefix(aux0, lane, e0, &i0);
    hist8_lane_prefix(aux1, lane, e1, &i1);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y0 = __shfl_up_sync(0xFFFFFFFFu, i0, o), y1 = __shfl_up_sync(0xFFFFFFFFu, i1, o);
        if (lane >= o) { i0 += y0; i1 += y1; }
    }
    const uint32_t b0 = lo0 + i0 - e0[PV - 1], b1 = lo1 + i1 - e1[PV - 1];
#pragma unroll
    for (int j = 0; j < PV; j++) { r0[j] = b0 + e0[j]; r1[j] = b1 + e1[j]; }
    __syncwarp();
    return true;
}

// r[j] = seg_begin + #(A <= qe[j]) for a chunk of a START-sorted probe table.  Interval ends are only sorted when all
// probes have the same length.  When they are (or when nothing is sorted) this is warp_window_ranks keyed on the ends.
// Sorted starts with unsorted ends -- variable-length probes, the usual case of interval tables that are not reads -- get a
// FORWARD search: the window that holds every rank of the chunk starts at the cursor (everything in front of it is <= the
// smallest start, hence <= every end), is staged in shared memory round by round until an element exceeds the largest end
// (typically a few dozen elements, whatever the probe lengths), and every row binary-searches its end in it with a
// warp-uniform trip count.  (The first version searched the starts inversely and let every row walk forward over the
// elements inside its probe: 0.86 ms per 125M probes of 100-200 bp and 1.58 ms at 1-3000 bp, against 0.47 ms for fixed
// lengths; the forward search: 0.65 and 0.69 ms, nearest 0.685 -> 0.568 ms.)  Degenerate rows (qe < qs) ask the directory.
// Same contract as warp_window_ranks.
// The forward search itself: r[j] = global index of the first element of the key's segment that is > q[j] (LT: >= q[j]).
// mn / mx = smallest / largest START of the (start-sorted, full) chunk: they steer the cursor; mxq = largest query, >= mx.
template <bool LT, typename CUR>
__device__ __forceinline__ bool warp_forward_ranks(CUR& cur, const DirView& d, const KeyDir& k, uint32_t kid,
                                                   const int32_t (&q)[PV], int32_t mn, int32_t mx, int32_t mxq, uint32_t* aux,
                                                   uint32_t (&r)[PV]) {
    const int lane = threadIdx.x & 31;
    const bool reuse = cur.valid && cur.key == kid && mn >= cur.bound;
    const uint32_t lo = reuse ? cur.pos : (LT ? rank_lt_kd(d, k, mn) : rank_le_kd(d, k, mn));
    int32_t* w = reinterpret_cast<int32_t*>(aux);
    __syncwarp();
    uint32_t wn_q = 0, wn_c = 0, staged = 0;
    bool open = true;
#pragma unroll 1
    for (; open && staged < (uint32_t)TP_WTILE; staged += 32) {
        const uint32_t idx = lo + staged + lane;
        int32_t v;
        if (reuse && staged == 0) v = cur.pre;  // loaded while the previous chunk was being finished
        else v = idx < k.seg_end ? __ldg(d.A + idx) : INT32_MAX;
        w[staged + lane] = v;
        const bool have = idx < k.seg_end;
        const uint32_t mq = __ballot_sync(0xFFFFFFFFu, have && (LT ? v < mxq : v <= mxq));
        wn_q += __popc(mq);
        wn_c += __popc(__ballot_sync(0xFFFFFFFFu, have && (LT ? v < mx : v <= mx)));
        open = (mq == 0xFFFFFFFFu);
    }
    __syncwarp();
    if (open) {  // more than 256 build values under one chunk: not this kernel's case
        cur.valid = false; cur.wide = true;
        return false;
    }
    cur.pos = lo + wn_c; cur.bound = mx; cur.key = kid; cur.valid = true;
    cur.prefetch(d.A, k.seg_end, lane);
    uint32_t pos[PV];
#pragma unroll
    for (int j = 0; j < PV; j++) pos[j] = 0;
    uint32_t len = wn_q;  // branchless bound search with a warp-uniform trip count; the 8 rows of a lane advance in lock step
#pragma unroll 1
    while (len > 1) {
        const uint32_t half = len >> 1;
#pragma unroll
        for (int j = 0; j < PV; j++) {
            const int32_t v = w[pos[j] + half - 1];
            pos[j] += (LT ? v < q[j] : v <= q[j]) ? half : 0u;
        }
        len -= half;
    }
#pragma unroll
    for (int j = 0; j < PV; j++) {
        if (wn_q) { const int32_t v = w[pos[j]]; pos[j] += (LT ? v < q[j] : v <= q[j]) ? 1u : 0u; }
        r[j] = lo + pos[j];
    }
    __syncwarp();
    return true;
}

template <typename CUR>
__device__ __forceinline__ bool warp_ranks_le_end(CUR& cur, const DirView& d, const KeyDir& k, uint32_t kid,
                                                  const int32_t (&qs)[PV], const int32_t (&qe)[PV], bool active, bool mono_s,
                                                  bool mono_e, const int32_t* Qs, const int32_t* Qe, uint32_t* aux,
                                                  uint32_t (&r)[PV]) {
    if (mono_e || !mono_s) return warp_window_ranks<false>(cur, d, k, kid, qe, active, mono_e, Qe, aux, r);
    const int32_t mn = __shfl_sync(0xFFFFFFFFu, qs[0], 0), mx = __shfl_sync(0xFFFFFFFFu, qs[PV - 1], 31);
    int32_t mxe = qe[0];
#pragma unroll
    for (int j = 1; j < PV; j++) mxe = max(mxe, qe[j]);
    mxe = max(__reduce_max_sync(0xFFFFFFFFu, mxe), mx);  // (a degenerate last row: the cursor still has to pass the starts)
    if (!warp_forward_ranks<false>(cur, d, k, kid, qe, mn, mx, mxe, aux, r)) return false;
#pragma unroll
    for (int j = 0; j < PV; j++)
        if (qe[j] < qs[j]) r[j] = rank_le_kd(d, k, qe[j]);  // degenerate row: its end may lie in front of the window
    return true;
}
// Transposed chunk stores (full 256-row chunks).  A lane owns 8 consecutive rows, so stored directly its outputs are 32 or
// 64 contiguous bytes and every store instruction of the warp touches 32 half-written sectors.  Staged through the warp's
// (now idle) histogram scratch, every store instruction writes 512 contiguous bytes: count 0.574 -> 0.549 ms per 125M probes.
// The staging area uses the padded layout of the histograms (16-byte unit u lives at u + u / 8): conflict-free both ways.
// x: one aux buffer (32-bit values) / two consecutive ones (64-bit values); out: the chunk's first output row.
__device__ __forceinline__ uint32_t pad_unit(uint32_t u) { return u + (u >> 3); }
__device__ __forceinline__ void store_chunk_i64_from32(uint32_t* x, int64_t* out, const uint32_t (&v)[PV], bool sign_extend) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    uint4* x4 = hist_lane_bins(x, lane);
    x4[0] = make_uint4(v[0], v[1], v[2], v[3]);
    x4[1] = make_uint4(v[4], v[5], v[6], v[7]);
    __syncwarp();
    longlong2* o2 = reinterpret_cast<longlong2*>(out) + lane;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint2 w = *reinterpret_cast<const uint2*>(x + hist_slot(2u * (k * 32 + lane)));
        __stcs(o2 + k * 32, sign_extend ? make_longlong2((long long)(int32_t)w.x, (long long)(int32_t)w.y)
                                        : make_longlong2((long long)w.x, (long long)w.y));
    }
    __syncwarp();
}
// the second half of store_chunk_i64_from32: the 256 32-bit values are already staged (value of row r at word hist_slot(r))
__device__ __forceinline__ void store_chunk_i64_staged32(uint32_t* x, int64_t* out, bool sign_extend) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    longlong2* o2 = reinterpret_cast<longlong2*>(out) + lane;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint2 w = *reinterpret_cast<const uint2*>(x + hist_slot(2u * (k * 32 + lane)));
        __stcs(o2 + k * 32, sign_extend ? make_longlong2((long long)(int32_t)w.x, (long long)(int32_t)w.y)
                                        : make_longlong2((long long)w.x, (long long)w.y));
    }
    __syncwarp();
}
__device__ __forceinline__ void store_chunk_i64(uint32_t* x, int64_t* out, const long long (&v)[PV]) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    longlong2* xl = reinterpret_cast<longlong2*>(x);
#pragma unroll
    for (int j = 0; j < PV / 2; j++) xl[pad_unit(lane * (PV / 2) + j)] = make_longlong2(v[2 * j], v[2 * j + 1]);
    __syncwarp();
    longlong2* o2 = reinterpret_cast<longlong2*>(out) + lane;
#pragma unroll
    for (int k = 0; k < 4; k++) __stcs(o2 + k * 32, xl[pad_unit(k * 32 + lane)]);
    __syncwarp();
}
__device__ __forceinline__ void store_chunk_u32(uint32_t* x, uint32_t* out, const uint32_t (&v)[PV]) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    uint4* x4 = reinterpret_cast<uint4*>(x);
    x4[pad_unit(2 * lane)] = make_uint4(v[0], v[1], v[2], v[3]);
    x4[pad_unit(2 * lane + 1)] = make_uint4(v[4], v[5], v[6], v[7]);
    __syncwarp();
    uint4* o4 = reinterpret_cast<uint4*>(out) + lane;
    __stcs(o4, x4[pad_unit(lane)]);
    __stcs(o4 + 32, x4[pad_unit(32 + lane)]);
    __syncwarp();
}
#endif

}  // namespace brg
