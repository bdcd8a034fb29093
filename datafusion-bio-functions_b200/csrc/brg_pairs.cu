// brg_pairs.cu -- K4: overlap pairs as count -> exclusive scan -> emit (sm_100a).
//
// Replaces IntervalJoinAlgorithm::get (Coitrees arm, interval_join.rs:849-863: every build row with
// b.start <= q.end && b.end >= q.start, closed, no guard for degenerate probes) and the RLE expansion of the
// probe-side indices (interval_join.rs:1616-1653).  Output is two Arrow UInt32 index buffers
// (build_idx, probe_idx), probe-major; matches of one probe come out in (start, end, position) order.
//
// Candidate range of a probe in the start-sorted rows of its key: [i0, n_le) with
//   n_le = #(start <= q.end)  and  i0 = first row whose prefix_max_end >= q.start
// (rows before i0 end before q.start, rows from n_le on start after q.end); inside it only `end >= q.start`
// has to be tested.  Both passes walk the same range, so the counts are exact for any input.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "brg_internal.cuh"
#include "brg_tile.cuh"

namespace brg {

constexpr int PR_THREADS = 256;

__device__ __forceinline__ void pair_range(const IndexView& ix, uint32_t c, int32_t qs, int32_t qe, uint32_t* i0,
                                           uint32_t* i1) {
    *i0 = *i1 = 0;
    if (c >= ix.n_keys) return;  // key absent from the build side (interval_join.rs:857)
    KeyDir ks = load_keydir(ix.d_sstart.kd, c);
    if (ks.seg_begin == ks.seg_end) return;
    uint32_t n_le = rank_le_kd(ix.d_sstart, ks, qe);
    if (n_le == ks.seg_begin) return;
    if (__ldg(ix.pme + n_le - 1) < qs) return;
    KeyDir kp = load_keydir(ix.d_pme.kd, c);
    *i0 = rank_lt_kd(ix.d_pme, kp, qs);
    *i1 = n_le;
}

__global__ void __launch_bounds__(PR_THREADS) pairs_count_kernel(IndexView ix, const uint32_t* __restrict__ key,
                                                                 const int32_t* __restrict__ start,
                                                                 const int32_t* __restrict__ end, uint64_t n,
                                                                 uint32_t* __restrict__ counts) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t qs = ld_stream(start + i), qe = ld_stream(end + i);
        uint32_t c = key ? ld_stream(key + i) : 0u;
        uint32_t i0, i1;
        pair_range(ix, c, qs, qe, &i0, &i1);
        uint32_t cnt = 0;
        for (uint32_t j = i0; j < i1; j++) cnt += (__ldg(ix.s_end + j) >= qs) ? 1u : 0u;
        counts[i] = cnt;
    }
}

__global__ void __launch_bounds__(PR_THREADS) pairs_emit_kernel(IndexView ix, const uint32_t* __restrict__ key,
                                                                const int32_t* __restrict__ start,
                                                                const int32_t* __restrict__ end, uint64_t n,
                                                                const uint64_t* __restrict__ offsets,
                                                                uint32_t* __restrict__ build_idx,
                                                                uint32_t* __restrict__ probe_idx) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t o = offsets[i];
        if (offsets[i + 1] == o) continue;  // nothing to write: skip the searches
        int32_t qs = ld_stream(start + i), qe = ld_stream(end + i);
        uint32_t c = key ? ld_stream(key + i) : 0u;
        uint32_t i0, i1;
        pair_range(ix, c, qs, qe, &i0, &i1);
        for (uint32_t j = i0; j < i1; j++) {
            if (__ldg(ix.s_end + j) >= qs) {
                build_idx[o] = __ldg(ix.s_pos + j);
                probe_idx[o] = (uint32_t)i;
                o++;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------
// Chunked pair pipeline (the fast path of BRG_PAIRS_INNER): the unit of the count -> scan -> emit scheme is a
// 256-row probe chunk, not a row.
//   pass 1  pairs_total_kernel_tiled : per chunk, sum of  #(start <= qe) - #(end < qs)  (the count_overlaps identity,
//           exact for build rows with start <= end and probes with qs <= qe; other rows walk their candidate range)
//   scan    exclusive scan of the n/256 chunk totals (tiny)
//   pass 2  pairs_emit_kernel_tiled  : recomputes the per-row counts of the chunk, prefix-sums them inside the warp,
//           walks each row's candidate range [first prefix_max_end >= qs, #(start <= qe)) and STAGES the matches in
//           shared memory so that the chunk leaves the SM as full, coalesced runs of build_idx / probe_idx
//           (per-row direct stores touch every 32 B sector up to eight times -- the old emit kernel's limit).
// Both passes read the probe columns through the per-warp TMA rings of brg_tile.cuh and find all ranks of a chunk with
// inverse searches keyed on the (sorted) probe starts.  Chunks that do not fit the pattern go to one-row-per-thread
// kernels through the to-do list; a chunk is always handled by exactly one of the two in each pass.
// ---------------------------------------------------------------------------------------------------
constexpr int PT_THREADS = 128;  // emit pass: 4 warps per CTA, 3 CTAs per SM (shared-memory bound)
constexpr int PT_WARPS = PT_THREADS / 32;
constexpr int PT_CAP = 1536;     // staged output slots per warp and round
struct __align__(16) PairStage {
    uint32_t b[PT_CAP];  // build row ids
    uint8_t p[PT_CAP];   // probe row within the chunk
};
struct __align__(128) PairSmem {
    WarpRing w[PT_WARPS];
    PairStage st[PT_WARPS];
};

// matches of one probe inside its candidate range (degenerate probes / generic fallback)
__device__ __forceinline__ uint32_t walk_count(const IndexView& ix, uint32_t i0, uint32_t i1, int32_t qs) {
    uint32_t c = 0;
    for (uint32_t i = i0; i < i1; i++) c += (__ldg(ix.s_end + i) >= qs) ? 1u : 0u;
    return c;
}

// ranks of one full chunk; false = defer the chunk
struct PairCursors {
    WarpCursor cs, cp, ce;
    KeyDir ks, kp, ke;
    uint32_t kd_key;
    __device__ __forceinline__ void reset() { cs.reset(); cp.reset(); ce.reset(); kd_key = 0xFFFFFFFFu; }
    __device__ __forceinline__ void invalidate() { cs.valid = false; cp.valid = false; ce.valid = false; }
};

template <bool NEED_I0>
__device__ __forceinline__ bool pair_chunk_ranks(const IndexView& ix, PairCursors& pc, const ProbeTile& t, uint32_t* aux0,
                                                 uint32_t* aux1, uint32_t* aux2, uint32_t (&n_le)[PV], uint32_t (&i0)[PV],
                                                 uint32_t (&cnt)[PV]) {
    uint32_t kuni;
    if (!warp_uniform_key(t.key, t.active, ix.n_keys, &kuni)) return false;
    if (kuni != pc.kd_key) {
        pc.kd_key = kuni;
        pc.ks = load_keydir(ix.d_sstart.kd, kuni);
        pc.kp = load_keydir(ix.d_pme.kd, kuni);
        pc.ke = load_keydir(ix.d_eend.kd, kuni);
    }
    const bool mono_s = chunk_mono(t.qs, t.full), mono_e = chunk_mono(t.qe, t.full);
    uint32_t ended[PV];
    if (!warp_ranks_le_end(pc.cs, ix.d_sstart, pc.ks, kuni, t.qs, t.qe, t.active, mono_s, mono_e, t.s_qs, t.s_qe, aux0, n_le))
        return false;
    if (!warp_window_ranks<true>(pc.ce, ix.d_eend, pc.ke, kuni, t.qs, t.active, mono_s, t.s_qs, aux1, ended)) return false;
    bool degenerate = false;
#pragma unroll
    for (int j = 0; j < PV; j++) degenerate |= t.qe[j] < t.qs[j];
    const bool any_deg = __any_sync(0xFFFFFFFFu, degenerate);
    if (NEED_I0 || any_deg) {
        if (!warp_window_ranks<true>(pc.cp, ix.d_pme, pc.kp, kuni, t.qs, t.active, mono_s, t.s_qs, aux2, i0)) return false;
    }
#pragma unroll
    for (int j = 0; j < PV; j++) {
        // start <= end on the build side (checked at build time): every row that ends before qs also starts before qe
        cnt[j] = n_le[j] - ended[j];
        if (t.qe[j] < t.qs[j]) cnt[j] = (i0[j] < n_le[j]) ? walk_count(ix, i0[j], n_le[j], t.qs[j]) : 0u;  // unguarded closed predicate
    }
    return true;
}

__global__ void __launch_bounds__(TP_THREADS, 2) pairs_total_kernel_tiled(IndexView ix, const uint32_t* __restrict__ key,
                                                                          const int32_t* __restrict__ start,
                                                                          const int32_t* __restrict__ end, uint64_t n_rows,
                                                                          uint32_t* __restrict__ chunk_tot, TodoList todo) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ProbeRing* ring = reinterpret_cast<ProbeRing*>(smem_raw);
    uint32_t* aux0 = ring->w[threadIdx.x >> 5].aux[0];
    uint32_t* aux1 = ring->w[threadIdx.x >> 5].aux[1];
    uint32_t* aux2 = ring->w[threadIdx.x >> 5].aux[2];
    ProbePipe pipe(ring, key, start, end, n_rows);  // n_rows is a multiple of 256: every chunk is full
    ProbeTile t;
    PairCursors pc;
    pc.reset();
    uint32_t fails = 0;
    while (pipe.next(t)) {
        uint32_t n_le[PV], i0[PV], cnt[PV];
        if (pair_chunk_ranks<false>(ix, pc, t, aux0, aux1, aux2, n_le, i0, cnt)) {
            uint32_t tot = 0;
#pragma unroll
            for (int j = 0; j < PV; j++) tot += cnt[j];
            tot = __reduce_add_sync(0xFFFFFFFFu, tot);
            if ((threadIdx.x & 31) == 0) chunk_tot[t.chunk] = tot;
            fails = 0;
        } else {
            pc.invalidate();
            todo_push(todo, t.chunk);
            fails++;
        }
        if (fails >= TP_BAIL) { pipe.bail_out(todo); break; }
        pipe.release();
    }
}

// writes the matches of the PV rows a lane owns, `o` = first output slot of the lane
template <bool DIRECT>
__device__ __forceinline__ void pair_walk(const IndexView& ix, const uint32_t (&i0)[PV], const uint32_t (&cnt)[PV],
                                          const int32_t (&qs)[PV], uint32_t o, PairStage& stg, uint32_t* gb, uint32_t* gp,
                                          uint32_t row0) {
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < PV; j++) {
        uint32_t left = cnt[j];
        uint32_t i = i0[j];
        while (left) {
            if (__ldg(ix.s_end + i) >= qs[j]) {
                const uint32_t pos = __ldg(ix.s_pos + i);
                if (DIRECT) { gb[o] = pos; gp[o] = row0 + lane * PV + j; }
                else { stg.b[o] = pos; stg.p[o] = (uint8_t)(lane * PV + j); }
                o++;
                left--;
            }
            i++;
        }
    }
}

__global__ void __launch_bounds__(PT_THREADS, 3) pairs_emit_kernel_tiled(IndexView ix, const uint32_t* __restrict__ key,
                                                                         const int32_t* __restrict__ start,
                                                                         const int32_t* __restrict__ end, uint64_t n_rows,
                                                                         const uint64_t* __restrict__ chunk_base,
                                                                         uint32_t* __restrict__ build_idx,
                                                                         uint32_t* __restrict__ probe_idx, TodoList todo) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PairSmem* sm = reinterpret_cast<PairSmem*>(smem_raw);
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t* aux0 = sm->w[warp].aux[0];
    uint32_t* aux1 = sm->w[warp].aux[1];
    uint32_t* aux2 = sm->w[warp].aux[2];
    PairStage& stg = sm->st[warp];
    ProbePipe pipe(sm->w, key, start, end, n_rows);  // full chunks only
    ProbeTile t;
    PairCursors pc;
    pc.reset();
    uint32_t fails = 0;
    while (pipe.next(t)) {
        uint32_t n_le[PV], i0[PV], cnt[PV];
        if (pair_chunk_ranks<true>(ix, pc, t, aux0, aux1, aux2, n_le, i0, cnt)) {
            fails = 0;
            uint32_t lane_tot = 0;
#pragma unroll
            for (int j = 0; j < PV; j++) lane_tot += cnt[j];
            uint32_t incl = lane_tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= (uint32_t)o) incl += y;
            }
            const uint32_t lane_off = incl - lane_tot;
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (total) {
                const uint64_t base = chunk_base[t.chunk];
                uint32_t* gb = build_idx + base;
                uint32_t* gp = probe_idx + base;
                const uint32_t row0 = (uint32_t)(t.chunk * TP_WTILE);
                uint32_t first = 0;  // lanes [0, first) are done (warp-uniform)
                while (first < 32) {
                    const uint32_t rb = __shfl_sync(0xFFFFFFFFu, lane_off, first);  // first slot of this round
                    const bool over = lane >= first && (incl - rb) > (uint32_t)PT_CAP;
                    const uint32_t nf = __ballot_sync(0xFFFFFFFFu, over);
                    const uint32_t stop = nf ? (uint32_t)(__ffs(nf) - 1) : 32u;     // first lane that no longer fits
                    if (stop == first) {  // one lane alone exceeds the stage: it writes straight to global memory
                        if (lane == first) pair_walk<true>(ix, i0, cnt, t.qs, lane_off, stg, gb, gp, row0);
                        first++;
                        continue;
                    }
                    const uint32_t next_off = __shfl_sync(0xFFFFFFFFu, lane_off, stop & 31u);
                    const uint32_t rend = (stop == 32u) ? total : next_off;
                    if (lane >= first && lane < stop) pair_walk<false>(ix, i0, cnt, t.qs, lane_off - rb, stg, gb, gp, row0);
                    __syncwarp();
                    for (uint32_t u = lane; u < rend - rb; u += 32) {
                        __stcs(gb + rb + u, stg.b[u]);
                        __stcs(gp + rb + u, row0 + stg.p[u]);
                    }
                    __syncwarp();
                    first = stop;
                }
            }
        } else {
            pc.invalidate();
            todo_push(todo, t.chunk);
            fails++;
        }
        if (fails >= TP_BAIL) { pipe.bail_out(todo); break; }
        pipe.release();
    }
}

// deferred chunks (and the partial last chunk, and everything when the tiled kernels cannot run): one row per thread
template <bool EMIT>
__global__ void __launch_bounds__(TP_WTILE) pairs_chunk_todo_kernel(IndexView ix, const uint32_t* __restrict__ key,
                                                                    const int32_t* __restrict__ start,
                                                                    const int32_t* __restrict__ end, uint64_t n_rows,
                                                                    uint32_t* __restrict__ chunk_tot,
                                                                    const uint64_t* __restrict__ chunk_base,
                                                                    uint32_t* __restrict__ build_idx,
                                                                    uint32_t* __restrict__ probe_idx, TodoList todo,
                                                                    uint32_t all_chunks, uint32_t tail_chunk) {
    __shared__ uint32_t wsum[TP_WTILE / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_list = todo.count ? *todo.count : all_chunks;
    const uint32_t n_items = n_list + (tail_chunk != 0xFFFFFFFFu ? 1u : 0u);
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t chunk = item < n_list ? (todo.count ? todo.chunks[item] : item) : tail_chunk;
        const uint64_t i = (uint64_t)chunk * TP_WTILE + threadIdx.x;
        uint32_t i0 = 0, i1 = 0, cnt = 0;
        int32_t qs = 0;
        if (i < n_rows) {
            qs = ld_stream(start + i);
            const int32_t qe = ld_stream(end + i);
            const uint32_t c = key ? ld_stream(key + i) : 0u;
            pair_range(ix, c, qs, qe, &i0, &i1);
            cnt = walk_count(ix, i0, i1, qs);
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < TP_WTILE / 32; w++) {
            const uint32_t v = wsum[w];
            before += (uint32_t)w < warp ? v : 0u;
            total += v;
        }
        __syncthreads();
        if (!EMIT) {
            if (threadIdx.x == 0) chunk_tot[chunk] = total;
        } else if (cnt) {
            uint64_t o = chunk_base[chunk] + before + incl - cnt;
            for (uint32_t j = i0; j < i1; j++) {
                if (__ldg(ix.s_end + j) >= qs) {
                    build_idx[o] = __ldg(ix.s_pos + j);
                    probe_idx[o] = (uint32_t)i;
                    o++;
                }
            }
        }
    }
}

// RightSemi / RightAnti (interval_join.rs:1433-1463): keep probe rows with (no) match
__global__ void semi_anti_flags_kernel(const uint32_t* __restrict__ counts, uint64_t n, int anti,
                                       uint32_t* __restrict__ flags) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool has = counts[i] != 0;
    flags[i] = (anti ? !has : has) ? 1u : 0u;
}
__global__ void compact_rows_kernel(const uint32_t* __restrict__ flags, const uint64_t* __restrict__ offsets, uint64_t n,
                                    uint32_t* __restrict__ probe_idx) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) probe_idx[offsets[i]] = (uint32_t)i;
}

static inline unsigned pairs_grid(uint64_t n) {
    uint64_t want = (n + PR_THREADS - 1) / PR_THREADS;
    uint64_t cap = (uint64_t)kSMs * 32;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

int launch_pairs_count(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                       uint32_t* counts, cudaStream_t st) {
    if (n == 0) return 0;
    pairs_count_kernel<<<pairs_grid(n), PR_THREADS, 0, st>>>(ix, key, s, e, n, counts);
    BRG_LAUNCH_CHECK();
    return 0;
}
int launch_pairs_emit(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                      const uint64_t* offsets, uint32_t* build_idx, uint32_t* probe_idx, cudaStream_t st) {
    if (n == 0) return 0;
    pairs_emit_kernel<<<pairs_grid(n), PR_THREADS, 0, st>>>(ix, key, s, e, n, offsets, build_idx, probe_idx);
    BRG_LAUNCH_CHECK();
    return 0;
}
int launch_semi_anti_flags(const uint32_t* counts, uint64_t n, int anti, uint32_t* flags, cudaStream_t st) {
    if (n == 0) return 0;
    semi_anti_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(counts, n, anti, flags);
    BRG_LAUNCH_CHECK();
    return 0;
}
int launch_compact_rows(const uint32_t* flags, const uint64_t* offsets, uint64_t n, uint32_t* probe_idx,
                        cudaStream_t st) {
    if (n == 0) return 0;
    compact_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(flags, offsets, n, probe_idx);
    BRG_LAUNCH_CHECK();
    return 0;
}


// ---- chunked pipeline launchers -----------------------------------------------------------------------------------
static bool pairs_tiled_enabled() {
    static const char* env = getenv("BRG_TILED");
    return !env || strstr(env, "pairs") != nullptr;
}
static inline bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline unsigned chunk_todo_grid(uint64_t chunks) {
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(chunks, (uint64_t)kSMs * 8));
}
struct ChunkPlan {
    uint64_t n_chunks, full_rows;
    uint32_t tail_chunk;
    bool tiled;
};
static ChunkPlan plan_chunks(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n) {
    ChunkPlan p;
    p.n_chunks = (n + TP_WTILE - 1) / TP_WTILE;
    p.full_rows = (n / TP_WTILE) * TP_WTILE;
    p.tail_chunk = (n % TP_WTILE) ? (uint32_t)(p.n_chunks - 1) : 0xFFFFFFFFu;
    p.tiled = pairs_tiled_enabled() && p.full_rows > 0 && aligned16p(key) && aligned16p(s) && aligned16p(e) &&
              (ix.flags & BRG_IDX_INVERTED) == 0;
    return p;
}
static int todo_list_alloc(uint64_t chunks, TodoList* td, cudaStream_t st) {
    void* p = nullptr;
    BRG_CUDA(cudaMallocAsync(&p, (chunks + 1) * sizeof(uint32_t), st));
    td->count = reinterpret_cast<uint32_t*>(p);
    td->chunks = td->count + 1;
    BRG_CUDA(cudaMemsetAsync(td->count, 0, sizeof(uint32_t), st));
    return 0;
}

uint64_t pairs_chunk_count(uint64_t n) { return (n + TP_WTILE - 1) / TP_WTILE; }

int launch_pairs_chunk_offsets(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                               uint64_t* chunk_base, cudaStream_t st) {
    const ChunkPlan p = plan_chunks(ix, key, s, e, n);
    if (p.n_chunks == 0) { BRG_CUDA(cudaMemsetAsync(chunk_base, 0, sizeof(uint64_t), st)); return 0; }
    uint32_t* tot = nullptr;
    uint64_t* scratch = nullptr;
    BRG_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tot), p.n_chunks * sizeof(uint32_t), st));
    BRG_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), (scan_scratch_words(p.n_chunks) + 16) * 4, st));
    int rc = 0;
    if (p.tiled) {
        TodoList todo;
        if ((rc = todo_list_alloc(p.n_chunks, &todo, st))) return rc;
        BRG_CUDA(cudaFuncSetAttribute((const void*)pairs_total_kernel_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProbeRing)));
        const uint64_t tiles = p.full_rows / TP_TILE + 1;
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)kSMs * 2));
        pairs_total_kernel_tiled<<<grid, TP_THREADS, sizeof(ProbeRing), st>>>(ix, key, s, e, p.full_rows, tot, todo);
        BRG_LAUNCH_CHECK();
        pairs_chunk_todo_kernel<false><<<chunk_todo_grid(p.n_chunks), TP_WTILE, 0, st>>>(ix, key, s, e, n, tot, nullptr, nullptr, nullptr,
                                                                                     todo, 0u, p.tail_chunk);
        BRG_LAUNCH_CHECK();
        BRG_CUDA(cudaFreeAsync(todo.count, st));
    } else {
        pairs_chunk_todo_kernel<false><<<chunk_todo_grid(p.n_chunks), TP_WTILE, 0, st>>>(ix, key, s, e, n, tot, nullptr, nullptr, nullptr,
                                                                                     TodoList{nullptr, nullptr}, (uint32_t)p.n_chunks,
                                                                                     0xFFFFFFFFu);
        BRG_LAUNCH_CHECK();
    }
    if ((rc = exclusive_scan_u32_to_u64(tot, chunk_base, p.n_chunks, scratch, st))) return rc;
    BRG_CUDA(cudaFreeAsync(tot, st));
    BRG_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

int launch_pairs_fill(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                      const uint64_t* chunk_base, uint32_t* build_idx, uint32_t* probe_idx, cudaStream_t st) {
    const ChunkPlan p = plan_chunks(ix, key, s, e, n);
    if (p.n_chunks == 0) return 0;
    int rc = 0;
    if (p.tiled) {
        TodoList todo;
        if ((rc = todo_list_alloc(p.n_chunks, &todo, st))) return rc;
        BRG_CUDA(cudaFuncSetAttribute((const void*)pairs_emit_kernel_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairSmem)));
        const uint64_t per_cta = (uint64_t)PT_WARPS * TP_WTILE;
        const uint64_t tiles = p.full_rows / per_cta + 1;
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)kSMs * 3));
        pairs_emit_kernel_tiled<<<grid, PT_THREADS, sizeof(PairSmem), st>>>(ix, key, s, e, p.full_rows, chunk_base, build_idx, probe_idx, todo);
        BRG_LAUNCH_CHECK();
        pairs_chunk_todo_kernel<true><<<chunk_todo_grid(p.n_chunks), TP_WTILE, 0, st>>>(ix, key, s, e, n, nullptr, chunk_base, build_idx,
                                                                                    probe_idx, todo, 0u, p.tail_chunk);
        BRG_LAUNCH_CHECK();
        BRG_CUDA(cudaFreeAsync(todo.count, st));
    } else {
        pairs_chunk_todo_kernel<true><<<chunk_todo_grid(p.n_chunks), TP_WTILE, 0, st>>>(ix, key, s, e, n, nullptr, chunk_base, build_idx,
                                                                                    probe_idx, TodoList{nullptr, nullptr},
                                                                                    (uint32_t)p.n_chunks, 0xFFFFFFFFu);
        BRG_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace brg
