// brg_tile.cuh -- tile pipeline and warp-window rank search shared by the probe kernels (sm_100a).
//
// Data movement: the probe columns are streamed with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx,
// L2 evict-first) into a shared-memory ring, issued by one thread per CTA, so HBM requests stay in flight
// independently of what the warps are doing.  (The first version of the kernels issued per-thread 4-byte loads and
// was latency-bound: 29 % DRAM throughput with every warp stalled on its probe loads, profiles/r1_count_v0_ncu.txt.)
//
// Search: every thread owns PV consecutive probe rows, a warp owns 32*PV consecutive rows, a CTA walks a CONTIGUOUS
// range of tiles.  Genomic probe tables are coordinate-sorted (the reference relies on that locality too,
// nearest.rs:320-322), so the ranks of one warp's rows fall into a short window of the sorted build array, and the
// window of the warp's next tile starts where the previous one ended.  Per sorted array the warp keeps a cursor
// (WarpCursor); a tile costs one coalesced load of the next few dozen build values (which ARE the window), a ballot
// to find the window end, and then every lane merges its PV rows against the window in shared memory: binary search
// for its first row, short forward walks for the following rows.  No dependent global loads, no directory lookups on
// the hot path.  Anything that breaks the pattern -- key change inside the warp, queries going backwards, a window
// wider than the shared buffer (unsorted input) -- falls back to the rank directory (brg_internal.cuh), per warp or
// per row, so results are exact for any input order.
#pragma once
#include "brg_internal.cuh"

namespace brg {

constexpr int TP_THREADS = 256;
constexpr int TP_WARPS = TP_THREADS / 32;
constexpr int PV = 8;                       // consecutive probe rows per thread
constexpr int TP_WTILE = 32 * PV;           // 256 consecutive rows per warp chunk
constexpr int TP_TILE = TP_THREADS * PV;    // rows in flight per CTA step
constexpr int TP_STAGES = 2;
constexpr int TP_WCAP = 256;                // window capacity per warp and array (elements)
constexpr int TP_NWIN = 2;                  // window buffers per warp

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

// ---- shared-memory layout ----------------------------------------------------------------------------------------
// Every warp runs its own pipeline over its own contiguous range of 256-row chunks: a private 2-stage ring fed by
// TMA bulk copies that lane 0 issues, a private mbarrier per stage, private window buffers.  Warps never wait for
// each other (no __syncthreads in the main loop).
struct __align__(128) WarpRing {
    int32_t start[TP_STAGES][TP_WTILE];
    int32_t end[TP_STAGES][TP_WTILE];
    uint32_t key[TP_STAGES][TP_WTILE];
    int32_t win[TP_NWIN][TP_WCAP];
    uint64_t full[TP_STAGES];
    uint64_t pad_[14];
};
struct __align__(128) ProbeRing {
    WarpRing w[TP_WARPS];
};

struct ProbeTile {
    int32_t qs[PV], qe[PV];
    uint32_t key[PV];
    uint64_t row0;   // global row of this thread's first probe
    bool active;     // thread owns rows in this chunk
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

    __device__ __forceinline__ ProbePipe(ProbeRing* r, const uint32_t* k, const int32_t* s, const int32_t* e,
                                         uint64_t rows)
        : ring(&r->w[threadIdx.x >> 5]), key(k), start(s), end(e), n_rows(rows), it(0) {
        const uint64_t n_chunks = (rows + TP_WTILE - 1) / TP_WTILE;
        const uint64_t n_warps = (uint64_t)gridDim.x * TP_WARPS;
        const uint64_t wid = (uint64_t)blockIdx.x * TP_WARPS + (threadIdx.x >> 5);
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
        if (tile.active) {
            const int4* ps = reinterpret_cast<const int4*>(&ring->start[s][my]);
            const int4* pe = reinterpret_cast<const int4*>(&ring->end[s][my]);
            int4 a = ps[0], b = ps[1], c4 = pe[0], d = pe[1];
            tile.qs[0] = a.x; tile.qs[1] = a.y; tile.qs[2] = a.z; tile.qs[3] = a.w;
            tile.qs[4] = b.x; tile.qs[5] = b.y; tile.qs[6] = b.z; tile.qs[7] = b.w;
            tile.qe[0] = c4.x; tile.qe[1] = c4.y; tile.qe[2] = c4.z; tile.qe[3] = c4.w;
            tile.qe[4] = d.x; tile.qe[5] = d.y; tile.qe[6] = d.z; tile.qe[7] = d.w;
            if (key) {
                const uint4* pk = reinterpret_cast<const uint4*>(&ring->key[s][my]);
                uint4 k0 = pk[0], k1 = pk[1];
                tile.key[0] = k0.x; tile.key[1] = k0.y; tile.key[2] = k0.z; tile.key[3] = k0.w;
                tile.key[4] = k1.x; tile.key[5] = k1.y; tile.key[6] = k1.z; tile.key[7] = k1.w;
            } else {
#pragma unroll
                for (int j = 0; j < PV; j++) tile.key[j] = 0u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < PV; j++) { tile.qs[j] = 0; tile.qe[j] = 0; tile.key[j] = 0xFFFFFFFFu; }
        }
        __syncwarp();  // every lane holds its rows in registers: the stage can be refilled
        if ((threadIdx.x & 31) == 0) issue(it + TP_STAGES);
        it++;
        return true;
    }
};

// All PV keys of every lane of the warp equal one valid key?  (AND == OR over all 256 keys.)  Inactive lanes pass
// their own neutral values via `active`.
__device__ __forceinline__ bool warp_uniform_key(const uint32_t (&key)[PV], bool active, uint32_t n_keys, uint32_t* kout) {
    uint32_t a = 0xFFFFFFFFu, o = 0u;
    if (active) {
#pragma unroll
        for (int j = 0; j < PV; j++) { a &= key[j]; o |= key[j]; }
    }
    a = __reduce_and_sync(0xFFFFFFFFu, a);
    o = __reduce_or_sync(0xFFFFFFFFu, o);
    *kout = o;
    return a == o && o < n_keys;  // (no active lane: a = ~0, o = 0 -> false)
}

// ---- warp cursor over one sorted array ---------------------------------------------------------------------------
// Invariant while valid: every element of key `key`'s segment before `pos` is <= bound (LT: < bound), hence `pos` is a
// legal window start for any tile whose smallest query is >= bound.
struct WarpCursor {
    uint32_t pos;
    int32_t bound;
    uint32_t key;
    bool valid;
    __device__ __forceinline__ void reset() { valid = false; pos = 0; bound = 0; key = 0xFFFFFFFFu; }
};

// r[j] = global index of the first element of key k's segment that is  > q[j] (LT = false)  or  >= q[j] (LT = true).
// Must be called by all 32 lanes; `k`/`kid` warp-uniform.  Lanes with active == false contribute nothing and get
// unspecified ranks.  `win` is this warp's shared window buffer (TP_WCAP entries).
template <bool LT>
__device__ __forceinline__ void warp_window_ranks(WarpCursor& cur, const DirView& d, const KeyDir& k, uint32_t kid,
                                                  const int32_t (&q)[PV], bool active, uint32_t (&r)[PV], int32_t* win) {
    const int lane = threadIdx.x & 31;
    int32_t mn = INT32_MAX, mx = INT32_MIN;
    if (active) {
#pragma unroll
        for (int j = 0; j < PV; j++) { mn = min(mn, q[j]); mx = max(mx, q[j]); }
    }
    mn = __reduce_min_sync(0xFFFFFFFFu, mn);
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    if (mn > mx) return;  // no active lane in this warp
    // ---- window start: the cursor when the queries moved forward, else a directory lookup ------------------------
    uint32_t lo;
    if (cur.valid && cur.key == kid && mn >= cur.bound) lo = cur.pos;
    else lo = LT ? rank_lt_kd(d, k, mn) : rank_le_kd(d, k, mn);
    // ---- window = the next build values up to the largest query: coalesced chunk loads + ballots ----------------
    uint32_t wn = 0;
    bool open = true;  // the window may extend beyond what has been loaded so far
    constexpr int CH = 2;
#pragma unroll 1
    for (uint32_t base = 0; open && base + 32 * CH <= (uint32_t)TP_WCAP; base += 32 * CH) {
        int32_t v[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const uint32_t idx = lo + base + 32 * c + lane;
            v[c] = idx < k.seg_end ? __ldg(d.A + idx) : INT32_MAX;
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const uint32_t idx = lo + base + 32 * c + lane;
            const bool in = idx < k.seg_end && (LT ? v[c] < mx : v[c] <= mx);
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, in);
            win[base + 32 * c + lane] = in ? v[c] : INT32_MAX;
            if (open) {
                wn += __popc(m);
                open = (m == 0xFFFFFFFFu);
            }
        }
    }
    __syncwarp();
    if (!open) {
        cur.pos = lo + wn; cur.bound = mx; cur.key = kid; cur.valid = true;
        // wn < TP_WCAP and win[wn] is a +inf pad: merge every lane's rows against the window
        uint32_t p = 0;
        {   // first row: binary search over [0, wn)
            uint32_t a = 0, b = wn;
            const int32_t q0 = q[0];
            while (a < b) {
                const uint32_t mid = (a + b) >> 1;
                const int32_t v = win[mid];
                if (LT ? v < q0 : v <= q0) a = mid + 1; else b = mid;
            }
            p = a;
        }
        r[0] = lo + p;
#pragma unroll
        for (int j = 1; j < PV; j++) {
            const int32_t qj = q[j];
            if (qj >= q[j - 1]) {
                // rows are sorted: walk forward (a handful of steps), then finish by binary search if still short
                int guard = 0;
                while (p < wn && (LT ? win[p] < qj : win[p] <= qj)) {
                    p++;
                    if (++guard == 8) {
                        uint32_t a = p, b = wn;
                        while (a < b) {
                            const uint32_t mid = (a + b) >> 1;
                            const int32_t v = win[mid];
                            if (LT ? v < qj : v <= qj) a = mid + 1; else b = mid;
                        }
                        p = a;
                        break;
                    }
                }
            } else {
                uint32_t a = 0, b = p;
                while (a < b) {
                    const uint32_t mid = (a + b) >> 1;
                    const int32_t v = win[mid];
                    if (LT ? v < qj : v <= qj) a = mid + 1; else b = mid;
                }
                p = a;
            }
            r[j] = lo + p;
        }
        __syncwarp();
        return;
    }
    // window wider than the shared buffer (unsorted probes): per-row directory search
    cur.valid = false;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < PV; j++) r[j] = active ? (LT ? rank_lt_kd(d, k, q[j]) : rank_le_kd(d, k, q[j])) : k.seg_begin;
}
#endif

}  // namespace brg
