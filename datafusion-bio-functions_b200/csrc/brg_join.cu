// brg_join.cu -- K4 in ONE launch: the overlap-pair join as  window -> ranges -> count -> output slots -> emit  (sm_100a).
//
// Replaces IntervalJoinAlgorithm::get (Coitrees arm, interval_join.rs:849-863: every build row with
// b.start <= q.end && b.end >= q.start, closed, no guard for degenerate probes) and the RLE expansion of the probe-side
// indices (interval_join.rs:1616-1653), like the three-launch pipeline of brg_pairs.cu -- but the candidate ranges of the
// probe rows never leave the SM: nothing is written to HBM except the pairs themselves (ncu: 632 MB of DRAM traffic against
// 642 MB algorithmic on the 10M x 10M join; the three-launch form moves 860 MB).
//
// Unit of work: a chunk of 32 x PV probe rows (PV = 4: 128 rows), owned by ONE warp; lane l owns rows l, l + 32, l + 64, l + 96.
// Chunks are handed out in order by an atomic ticket; warps never synchronise with each other (no __syncthreads in the loop).
//
// Fast path (rows of one key, coordinate-sorted -- the shape of real interval tables):
//   1. the chunk's WINDOW of the start-sorted build rows starts at W = the first row in the prefix_max_end directory bin of
//      (smallest start of the chunk) - 1: every earlier row of the key ends before any row of the chunk starts.  ONE dependent
//      load; the window is the next FJ_WCAP rows, copied by TMA (start, prefix_max_end | end, row id on two mbarriers);
//   2. per row  i0 = first window row with prefix_max_end >= qs  and  n_le = #(window start <= qe): branch-free binary
//      searches over the window in shared memory, a fixed number of unrolled steps, no bound checks (sentinels behind the
//      window), brg_win.cuh; a row that reaches the end of a truncated window sends the chunk to the slow path;
//   3. the candidate slots [i0, n_le) of all rows are flattened in row order across the warp (a byte map slot -> row built with
//      one scattered store per non-empty row and a byte-wise prefix sum in registers), 32 slots are tested per step
//      (`end >= qs`, from shared memory), compacted with a ballot and STAGED in shared memory as (window row | row << 16)
//      words over the searched arrays, which are dead by then;
//   4. the chunk gets its output slots and the staged pairs leave, row ids resolved from the window, one 32-bit element per
//      lane and store (every store instruction writes 128 contiguous bytes).
// Output slots, two ways (flag BRG_JOIN_ORDERED of brg_pairs_join_dev):
//   any order (default): one atomicAdd of the chunk's pair count on a global cursor.  The pairs of a chunk are contiguous and
//     probe-major, the chunks land in the order they finish; the optional chunk-slot table says where.  The reference's row
//     order is unspecified as well (coitrees visit order inside a probe row, batch arrival order across partitions), and row
//     sets are compared after a canonical sort.
//   ordered: a two-level decoupled look-back over the chunks' counts, so that the output is probe-major throughout (the order
//     of brg_pairs.cu).  A chunk can only be written once EVERY earlier chunk is counted: with ~3500 chunks in flight each
//     chunk waits for the slowest of its predecessors, about as long as it works (measured 17k of 38k cycles per chunk; a
//     second staging buffer that defers the wait by one chunk made it worse, 0.79 vs 0.45 ms -- profiles/r2_join_experiments.txt).
// Slow path (key boundary inside the chunk, unsorted rows, window or candidate count above the staging capacity, partial
// last chunk): per-row rank-directory searches, the count from the count_overlaps identity (or a counting pass over the
// flattened slots), the output slots, and a pass that writes the pairs straight to global memory.  Exact for any input.
//
// Order inside a chunk: probe-major; the matches of one probe in (start, end, row id) order -- the same as brg_pairs.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "brg_internal.cuh"
#include "brg_tile.cuh"
#include "brg_win.cuh"

namespace brg {

#ifndef BRG_FJ_PV
#define BRG_FJ_PV 4             // rows per lane.  4 (128-row chunks) instead of the 8 of the tiled probe kernels: half the shared memory
#endif                          // and registers per warp = twice the warps (8 rows per lane: 0.254 ms against 0.190 ms)
constexpr int FJ_PV = BRG_FJ_PV;
constexpr int FJ_ROWS = 32 * FJ_PV;  // rows per chunk
#ifndef BRG_FJ_WCAP
#define BRG_FJ_WCAP (FJ_ROWS + FJ_ROWS / 2)   // window rows staged per chunk
#endif
#ifndef BRG_FJ_SLOTS
#define BRG_FJ_SLOTS (FJ_ROWS * 12)           // candidate slots flattened at a time (bytes of the slot -> row map)
#endif
#ifndef BRG_FJ_STAGE
#define BRG_FJ_STAGE (FJ_ROWS * 8)            // pairs staged per chunk (one word each)
#endif
#ifndef BRG_FJ_WARPS
#define BRG_FJ_WARPS 8          // warps per CTA
#endif
#ifndef BRG_FJ_CTAS
#define BRG_FJ_CTAS 3
#endif
#ifndef BRG_FJ_VEC
#define BRG_FJ_VEC 0            // 1: 128-bit stores of the staged pairs
#endif
constexpr int FJ_WCAP = BRG_FJ_WCAP, FJ_SLOTS = BRG_FJ_SLOTS, FJ_STAGE = BRG_FJ_STAGE, FJ_WARPS = BRG_FJ_WARPS;
#ifndef BRG_FJ_BATCH
#define BRG_FJ_BATCH 2
#endif
constexpr int FJ_BATCH = BRG_FJ_BATCH;  // chunks per ticket in any-order mode
#ifndef BRG_FJ_U
#define BRG_FJ_U 4
#endif
constexpr int FJ_U = BRG_FJ_U;  // slot steps in flight per iteration of the test loops
constexpr int FJ_RSLOTS = FJ_ROWS + 2;
#ifndef BRG_FJ_SPIN_NS
#define BRG_FJ_SPIN_NS 100
#endif
constexpr unsigned FJ_SPIN_NS = BRG_FJ_SPIN_NS;
constexpr int FJ_MAPQ = FJ_SLOTS / 512;  // 16-byte words of the slot -> row map per lane
constexpr int FJ_PAD = win_pad(FJ_WCAP);  // sentinel rows behind the window
constexpr int FJ_KEYS = 64;       // keys whose directory constants are staged per CTA
static_assert(FJ_PV == 4 || FJ_PV == 8, "two 16-bit candidate counts per packed warp scan");
static_assert(FJ_SLOTS % 512 == 0 && FJ_SLOTS <= 65536 && FJ_WCAP <= 32768 && FJ_WCAP % 4 == 0 && FJ_STAGE % 4 == 0 && FJ_ROWS <= 256,
              "packing of the row words");

// Lane l owns the rows l, l + 32, l + 64, ... of its chunk (not PV consecutive ones): in every step of the shared-memory searches
// the 32 lanes then work on 32 CONSECUTIVE sorted queries, whose positions in the window are about one row apart -- broadcasts
// and consecutive banks.  With consecutive rows per lane the lanes sat ~PV rows apart in the window: 4-way bank conflicts on every
// search step, 43 % of all shared-memory wavefronts of a kernel that is bound by the shared-memory pipe (profiles/r2_join_*).
__device__ __forceinline__ uint32_t fj_row(uint32_t lane, int j) { return lane + 32u * (uint32_t)j; }

struct __align__(16) JoinWarp {
    union {
        struct {
            int32_t start_[FJ_WCAP + 4 + FJ_PAD];  // window starts / prefix_max_end from [4] on (TMA); [3] is the guard word read at
            int32_t pme_[FJ_WCAP + 4 + FJ_PAD];    // index -1; INT32_MAX sentinels behind the window's last row
        } s;
        uint32_t stage[FJ_STAGE];     // the chunk's pairs as (window row | row << 16) -- the searched arrays are dead by then
        uint4 row4[FJ_RSLOTS];        // slow path: {exclusive prefix of the candidate counts, i0, qs, candidates}
    } u;
    uint32_t w_pos[FJ_WCAP];          // window: row ids of the build rows [Wa, Wa + nw) (TMA)
    int32_t w_end[FJ_WCAP];           // window: their ends (TMA)
    uint2 row2[FJ_RSLOTS];            // {qs, i0 - prefix} per row
    uint8_t rowof[FJ_SLOTS + 144];    // slot -> row (+ a zeroed tail for the last, partial step of the test loop)
    uint64_t bar[2];                  // mbarriers: window (start, prefix_max_end), window (end, position)
};
static_assert(sizeof(uint32_t) * FJ_STAGE >= sizeof(int32_t) * 2 * (FJ_WCAP + 4 + FJ_PAD) && sizeof(uint32_t) * FJ_STAGE >= sizeof(uint4) * FJ_RSLOTS,
              "the staging area covers the searched arrays and the slow path's row records");

// -DBRG_FJ_PROF: per-phase clock64() sums of the join kernel (profiles/exp_join.py --prof); not in the shipped library
#ifdef BRG_FJ_PROF
__device__ unsigned long long g_fj_prof[16];
#define FJ_PROF_DECL long long fj_t0 = clock64(); unsigned long long fj_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define FJ_T(i) { const long long fj_now = clock64(); fj_acc[i] += (unsigned long long)(fj_now - fj_t0); fj_t0 = fj_now; }
#define FJ_N(i) { fj_acc[i] += 1; }
#define FJ_PROF_FLUSH if (lane == 0) { _Pragma("unroll") for (int i = 0; i < 16; i++) atomicAdd(&g_fj_prof[i], fj_acc[i]); }
#else
#define FJ_PROF_DECL
#define FJ_T(i)
#define FJ_N(i)
#define FJ_PROF_FLUSH
#endif

constexpr uint64_t FJ_FLAG_AGG = 1ull << 62, FJ_FLAG_INC = 2ull << 62, FJ_VAL = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// ---- ordered mode: the scan over the chunks' pair counts -------------------------------------------------------------------
// A flat decoupled look-back does not work at this granularity: with thousands of warps in flight a chunk's nearest predecessor
// that knows its prefix is thousands of words back (measured: 55 dependent polls, 29 us per chunk).  So the scan has TWO levels;
// chunks are grouped by 32.  All words are 64 bits with the value inside the flagged word, so no fence is needed.
//   chunk word  cstate[c]: 0 -> (AGG | count); a chunk sums the words of the earlier chunks of its group with ONE poll;
//   group word  gstate[g]: 0 -> (AGG | group total) -> (INC | inclusive prefix), written by the group's last chunk; every chunk
//   runs the decoupled look-back over the GROUP words from g - 1 backwards (32 groups = 1024 chunks per poll).
// (Tried instead: a packed atomic per group so that whichever chunk finishes last publishes the group total, and readers publish
//  the inclusive prefixes -- 1.58 ms against 0.45 ms for this form.)
struct ScanState {
    uint64_t* cursor;  // any order: the next free output slot
    uint64_t* cstate;
    uint64_t* gstate;
    uint32_t n_chunks;
    int ordered;
};
__device__ __forceinline__ uint64_t chunk_lookback(const ScanState& sc, uint32_t chunk, uint64_t own, uint32_t lane) {
    const uint32_t g = chunk >> 5, idx = chunk & 31u;
    if (lane == 0) st_relaxed_u64(sc.cstate + chunk, FJ_FLAG_AGG | own);
    uint64_t v;
    for (;;) {
        v = lane < idx ? ld_relaxed_u64(sc.cstate + (g << 5) + lane) : FJ_FLAG_AGG;
        if (!__any_sync(0xFFFFFFFFu, (v >> 62) == 0)) break;
        __nanosleep(FJ_SPIN_NS);  // (a waiting warp should not take issue slots from the ones it waits for)
    }
    const uint64_t in_group = warp_sum_u64(lane < idx ? (v & FJ_VAL) : 0ull);
    const bool leader = idx == 31u || chunk == sc.n_chunks - 1;
    if (leader && g && lane == 0) st_relaxed_u64(sc.gstate + g, FJ_FLAG_AGG | (in_group + own));
    uint64_t gbase = 0;
    if (g) {
        int64_t j = (int64_t)g - 1 - (int64_t)lane;
        for (;;) {
            for (;;) {
                v = j >= 0 ? ld_relaxed_u64(sc.gstate + j) : FJ_FLAG_INC;  // before group 0: inclusive prefix 0
                if (!__any_sync(0xFFFFFFFFu, (v >> 62) == 0)) break;
                __nanosleep(FJ_SPIN_NS);
            }
            const uint32_t inc = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
            const uint32_t first = inc ? (uint32_t)__ffs(inc) - 1u : 31u;  // the closest earlier group that knows its prefix
            gbase += warp_sum_u64(lane <= first ? (v & FJ_VAL) : 0ull);
            if (inc) break;
            j -= 32;
        }
    }
    if (leader && lane == 0) st_relaxed_u64(sc.gstate + g, FJ_FLAG_INC | (gbase + in_group + own));
    return gbase + in_group;
}

// the first output slot of a counted chunk, either way
__device__ __forceinline__ uint64_t chunk_slots(const ScanState& sc, uint32_t chunk, uint64_t own, uint32_t lane) {
    if (!sc.ordered) {
        unsigned long long b = 0;
        if (lane == 0 && own) b = atomicAdd(reinterpret_cast<unsigned long long*>(sc.cursor), (unsigned long long)own);
        return __shfl_sync(0xFFFFFFFFu, b, 0);
    }
    return chunk_lookback(sc, chunk, own, lane);
}

// exclusive prefix of the rows' candidate counts IN ROW ORDER (row = lane + 32 j): one warp scan per group of 32 rows, two groups
// packed into one 32-bit scan (counts <= FJ_WCAP, sums per group < 65536).  Returns the chunk's total.
__device__ __forceinline__ uint32_t fj_row_prefix16(const uint32_t (&cand)[FJ_PV], uint32_t lane, uint32_t (&pj)[FJ_PV]) {
    uint32_t base = 0;
#pragma unroll
    for (int h = 0; h < FJ_PV / 2; h++) {
        const uint32_t own = cand[2 * h] | (cand[2 * h + 1] << 16);
        uint32_t incl = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const uint32_t excl = incl - own;
        pj[2 * h] = base + (excl & 0xFFFFu);
        base += tot & 0xFFFFu;
        pj[2 * h + 1] = base + (excl >> 16);
        base += tot >> 16;
    }
    return base;
}

// the slot -> row map of rows [ra, rb) whose candidates start at slot Pa
__device__ __forceinline__ void fj_build_rowof(uint8_t* rowof, uint32_t n_slots, uint32_t ra, uint32_t rb, uint32_t Pa,
                                               const uint32_t (&pj)[FJ_PV], const uint32_t (&cand)[FJ_PV], uint32_t lane) {
    // Row ids rise along the slots, so the map is a byte-wise PREFIX SUM of "row id minus the previous non-empty row's id" dropped at
    // every non-empty row's first slot: inside a 32-bit word that is one multiplication by 0x01010101 (ids are < 256: no carries
    // between the bytes).  (The first version took a byte-wise prefix MAX of the row ids with __vmaxu4 -- emulated on sm_100 with
    // ~8 instructions each, 450 of the kernel's 2300 instructions per chunk.)
    // The whole map (FJ_SLOTS bytes) is swept whatever n_slots is: every lane owns FJ_MAPQ consecutive 16-byte words of it and keeps
    // them in registers between one 128-bit load and one 128-bit store each (lane stride 16 * FJ_MAPQ bytes: conflict-free).  Slots
    // behind the last candidate inherit the last row: any valid row will do for the idle lanes of the last test step.
    uint4* mine = reinterpret_cast<uint4*>(rowof) + lane * FJ_MAPQ;
#pragma unroll
    for (int q = 0; q < FJ_MAPQ; q++) mine[q] = make_uint4(0u, 0u, 0u, 0u);
    // the previous non-empty row of each of the lane's rows, in row order: the highest set bit below the lane in the group's ballot,
    // else the last non-empty row of the groups before (0 before the first one, so that the map starts with that row's id itself)
    uint32_t prev[FJ_PV];
    uint32_t before_groups = 0;
    const uint32_t below = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < FJ_PV; j++) {
        const uint32_t r = fj_row(lane, j);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, r >= ra && r < rb && cand[j] != 0u);
        const uint32_t lower = m & below;
        prev[j] = lower ? 32u * (uint32_t)j + (31u - (uint32_t)__clz(lower)) : before_groups;
        if (m) before_groups = 32u * (uint32_t)j + (31u - (uint32_t)__clz(m));
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < FJ_PV; j++) {
        const uint32_t r = fj_row(lane, j);
        if (r >= ra && r < rb && cand[j]) rowof[pj[j] - Pa] = (uint8_t)(r - prev[j]);
    }
    __syncwarp();
    uint32_t v[4 * FJ_MAPQ];
#pragma unroll
    for (int q = 0; q < FJ_MAPQ; q++) {
        const uint4 t = mine[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    uint32_t carry = 0;  // sum of the lane's bytes so far
#pragma unroll
    for (int i = 0; i < 4 * FJ_MAPQ; i++) {
        v[i] = (v[i] + carry) * 0x01010101u;
        carry = v[i] >> 24;
    }
    uint32_t before = carry;  // exclusive prefix sum over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, before, o);
        if (lane >= (uint32_t)o) before += y;
    }
    before -= carry;
    const uint32_t bb = before * 0x01010101u;
#pragma unroll
    for (int q = 0; q < FJ_MAPQ; q++) mine[q] = make_uint4(v[4 * q] + bb, v[4 * q + 1] + bb, v[4 * q + 2] + bb, v[4 * q + 3] + bb);
    __syncwarp();
}

// Slow path, one pass over the flattened candidate slots of a chunk (rows and ranges in row4 and in the lane's registers,
// build rows read from global memory).  EMIT = false counts the pairs, EMIT = true writes them from slot `base` on.
template <bool EMIT>
__device__ __forceinline__ uint64_t fj_slow_pass(const IndexView& ix, JoinWarp& sm, const uint4* row4, const uint32_t (&pj)[FJ_PV],
                                                 const uint32_t (&cand)[FJ_PV], bool huge, uint32_t row0, uint64_t base,
                                                 uint64_t capacity, uint32_t* __restrict__ build_idx,
                                                 uint32_t* __restrict__ probe_idx, uint32_t lane) {
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint64_t run = 0;  // pairs of this chunk seen so far (warp-uniform)
    uint32_t ra = 0;   // rows [0, ra) are done (warp-uniform)
    while (ra < (uint32_t)FJ_ROWS) {
        const uint32_t Pa = row4[ra].x;
        uint32_t fit = 0;
        if (!huge) {  // rows [ra, rb) whose candidates fit the slot map together (their prefix is monotone)
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) fit += (fj_row(lane, j) >= ra && pj[j] + cand[j] - Pa <= (uint32_t)FJ_SLOTS) ? 1u : 0u;
            fit = __reduce_add_sync(0xFFFFFFFFu, fit);
        }
        if (fit == 0) {
            // one row alone exceeds the map (or the chunk's prefix does not fit 32 bits): the warp sweeps its range
            const uint4 rw = row4[ra];
            const uint32_t rI0 = rw.y, rN = rw.w;
            const int32_t rQ = (int32_t)rw.z;
            for (uint32_t sb = 0; sb < rN; sb += 32) {
                const uint32_t s = sb + lane;
                const bool m = s < rN && s >= sb && __ldg(ix.s_end + rI0 + s) >= rQ;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, m);
                if (EMIT && m) {
                    const uint64_t o = base + run + __popc(bal & lt_mask);
                    if (o < capacity) {
                        __stcs(build_idx + o, __ldg(ix.s_pos + rI0 + s));
                        __stcs(probe_idx + o, row0 + ra);
                    }
                }
                run += __popc(bal);
                if (sb + 32 < sb) break;  // (rN within 32 of 2^32)
            }
            ra++;
            continue;
        }
        const uint32_t rb = ra + fit;
        const uint32_t n_slots = row4[rb].x - Pa;  // (row 256 holds the chunk's total)
        if (n_slots) {
            fj_build_rowof(sm.rowof, n_slots, ra, rb, Pa, pj, cand, lane);
            for (uint32_t sb = 0; sb < n_slots; sb += 32 * FJ_U) {
                uint32_t rr[FJ_U], pos[FJ_U];
                bool mt[FJ_U];
#pragma unroll
                for (int u = 0; u < FJ_U; u++) {
                    const uint32_t s = sb + u * 32 + lane;
                    const bool v = s < n_slots;
                    rr[u] = v ? sm.rowof[s] : ra;
                    const uint4 rw = row4[rr[u]];
                    const uint32_t i = rw.y + (s + Pa - rw.x);
                    mt[u] = v && __ldg(ix.s_end + i) >= (int32_t)rw.z;
                    pos[u] = (EMIT && v) ? __ldg(ix.s_pos + i) : 0u;
                }
#pragma unroll
                for (int u = 0; u < FJ_U; u++) {
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, mt[u]);
                    if (EMIT && mt[u]) {
                        const uint64_t o = base + run + __popc(bal & lt_mask);
                        if (o < capacity) {
                            __stcs(build_idx + o, pos[u]);
                            __stcs(probe_idx + o, row0 + rr[u]);
                        }
                    }
                    run += __popc(bal);
                }
            }
            __syncwarp();  // the map is rebuilt by the next group
        }
        ra = rb;
    }
    return run;
}

// rank-directory search of the 8 rows of a lane in lock step: out[j] = global index of the first element of key c[j]'s
// segment that is > q[j]; rows with act[j] == false are skipped
__device__ __forceinline__ void fj_dir_ranks8(const DirView& d, const uint32_t (&c)[FJ_PV], const int32_t (&q)[FJ_PV],
                                              const bool (&act)[FJ_PV], uint32_t (&out)[FJ_PV]) {
    uint32_t lo[FJ_PV], hi[FJ_PV];
#pragma unroll
    for (int j = 0; j < FJ_PV; j++) {
        lo[j] = hi[j] = 0;
        if (act[j]) {
            const KeyDir k = load_keydir(d.kd, c[j]);
            rank_le_bin(d, k, q[j], &lo[j], &hi[j]);
        }
    }
    bool any = true;
    while (any) {
        any = false;
#pragma unroll
        for (int j = 0; j < FJ_PV; j++) {
            if (lo[j] < hi[j]) {
                const uint32_t mid = (lo[j] + hi[j]) >> 1;
                if (__ldg(d.A + mid) <= q[j]) lo[j] = mid + 1; else hi[j] = mid;
                any = any || lo[j] < hi[j];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < FJ_PV; j++) out[j] = lo[j];
}

__global__ void __launch_bounds__(FJ_WARPS * 32, BRG_FJ_CTAS)
pairs_join_kernel(IndexView ix, const uint32_t* __restrict__ key, const int32_t* __restrict__ start,
                  const int32_t* __restrict__ end, uint64_t n_rows, uint32_t* ticket, ScanState sc, uint64_t capacity,
                  uint32_t* __restrict__ build_idx, uint32_t* __restrict__ probe_idx, uint64_t* __restrict__ total_out,
                  uint64_t* __restrict__ chunk_slots_out, int vec_in, int vec_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ KeyDir skd[FJ_KEYS];  // the per-key constants of the prefix_max_end directory (a table has a few dozen keys)
    JoinWarp& sm = reinterpret_cast<JoinWarp*>(smem_raw)[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t n_chunks = sc.n_chunks;
    int32_t* const w_start = sm.u.s.start_ + 4;
    int32_t* const w_pme = sm.u.s.pme_ + 4;
    const bool keys_staged = ix.n_keys <= (uint32_t)FJ_KEYS;
    if (keys_staged) {
        const uint32_t words = ix.n_keys * (uint32_t)(sizeof(KeyDir) / 4);
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x)
            reinterpret_cast<uint32_t*>(skd)[i] = __ldg(reinterpret_cast<const uint32_t*>(ix.d_pme.kd) + i);
    }
    if (lane == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); mbar_fence_init(); }
    for (uint32_t x = lane; x < 144u; x += 32) sm.rowof[FJ_SLOTS + x] = 0;  // the map's tail: row 0, for the idle lanes of a last test step
    __syncthreads();
    uint32_t ph_a = 0, ph_b = 0;  // mbarrier phases: window (start, prefix_max_end), window (end, position)
    FJ_PROF_DECL
    // A warp claims a chunk when it is ready to start it, never earlier.  Ordered mode: a warp that WAITS in the look-back while it
    // holds a chunk it has not counted makes every later chunk wait for it -- claiming two tickets ahead serialised the whole
    // kernel (64 ms).  Any-order mode: fetching the next chunk's ticket and probe rows ahead costs 12 registers per thread and
    // spills at 80 (0.213 ms with, 0.196 ms without).
    auto load_chunk = [&](uint32_t c, int32_t (&oqs)[FJ_PV], int32_t (&oqe)[FJ_PV], uint32_t (&okk)[FJ_PV]) {
        const uint64_t r0 = (uint64_t)c * FJ_ROWS;
        if (r0 + FJ_ROWS <= n_rows) {  // (32 consecutive rows per load instruction: 128 contiguous bytes)
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) {
                const uint64_t r = r0 + fj_row(lane, j);
                oqs[j] = __ldcs(start + r);
                oqe[j] = __ldcs(end + r);
                okk[j] = key ? __ldcs(key + r) : 0u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) {
                const uint64_t r = r0 + fj_row(lane, j);
                const bool v = r < n_rows;
                oqs[j] = v ? __ldcs(start + r) : 0;
                oqe[j] = v ? __ldcs(end + r) : 0;
                okk[j] = v ? (key ? __ldcs(key + r) : 0u) : 0xFFFFFFFFu;  // rows past the end: no key
            }
        }
    };
    // Any-order mode claims FJ_BATCH consecutive chunks per ticket (a sparse join is bound by the chain of dependent round trips
    // per chunk -- ticket, probe rows, directory word, window, cursor: the batch saves the first for three chunks of four and lets
    // the probe rows of the whole batch be prefetched into L2 at once).  Ordered mode keeps one chunk per ticket: chunks of a batch
    // are counted one after the other, and everything behind them would wait that much longer.
    const uint32_t batch = sc.ordered ? 1u : (uint32_t)FJ_BATCH;
    uint32_t chunk = 0, batch_end = 0;
    for (;;) {
        if (chunk >= batch_end) {
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(ticket, batch);
            chunk = __shfl_sync(0xFFFFFFFFu, t, 0);
            if (chunk >= n_chunks) break;
            batch_end = min(chunk + batch, n_chunks);
            if (batch > 1u) {  // the batch's probe rows: (batch_end - chunk) * FJ_ROWS * 4 bytes per column, one 128-byte line per lane and step
                const uint64_t b0 = (uint64_t)chunk * FJ_ROWS, b1 = min((uint64_t)batch_end * FJ_ROWS, n_rows);
                for (uint64_t r = b0 + 32u * lane; r < b1; r += 32u * 32u) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(start + r));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(end + r));
                    if (key) asm volatile("prefetch.global.L2 [%0];" ::"l"(key + r));
                }
            }
        }
        const uint64_t row0 = (uint64_t)chunk * FJ_ROWS;
        int32_t qs[FJ_PV], qe[FJ_PV];
        uint32_t kk[FJ_PV];
        bool fast = row0 + FJ_ROWS <= n_rows;
        load_chunk(chunk, qs, qe, kk);
        uint64_t total = 0, base = 0;
        bool done = false;
        uint32_t n_staged = 0;
        FJ_T(0)
        if (fast) {
            // ---- one key, sorted starts? ---------------------------------------------------------------------------------
            const uint32_t k0 = __shfl_sync(0xFFFFFFFFu, kk[0], 0);
            bool ok = true;
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) ok = ok & (kk[j] == k0);
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) {  // row lane + 32 j against the next row: the next lane's, or lane 0's of the next group
                int32_t nx = __shfl_down_sync(0xFFFFFFFFu, qs[j], 1);
                if (j + 1 < FJ_PV) {
                    const int32_t wrap = __shfl_sync(0xFFFFFFFFu, qs[j + 1 < FJ_PV ? j + 1 : j], 0);
                    if (lane == 31) nx = wrap;
                } else if (lane == 31) {
                    nx = INT32_MAX;
                }
                ok = ok & (qs[j] <= nx);
            }
            fast = __all_sync(0xFFFFFFFFu, ok);
            if (fast && k0 >= ix.n_keys) done = true;  // key absent from the build side (interval_join.rs:857): no pairs
            if (fast && !done) {
                // ---- the chunk's window of the start-sorted build rows: FJ_WCAP rows from W on, by TMA ---------------------
                const int32_t qs_first = __shfl_sync(0xFFFFFFFFu, qs[0], 0);
                uint32_t W = 0, seg_begin = 0, seg_end = 0;
                if (lane == 0) {
                    const KeyDir k = keys_staged ? skd[k0] : load_keydir(ix.d_pme.kd, k0);
                    W = win_begin(ix.d_pme, k, qs_first);
                    seg_begin = k.seg_begin; seg_end = k.seg_end;
                }
                W = __shfl_sync(0xFFFFFFFFu, W, 0);
                seg_begin = __shfl_sync(0xFFFFFFFFu, seg_begin, 0);
                seg_end = __shfl_sync(0xFFFFFFFFu, seg_end, 0);
                const uint32_t Wa = W & ~3u;                                    // (TMA: 16-byte aligned source)
                const uint32_t nw = min((uint32_t)FJ_WCAP, seg_end - Wa);       // window rows [Wa, Wa + nw), searched from wlo on
                const uint32_t wlo = max(seg_begin, Wa) - Wa;                   // (0..3 rows of the previous key in front)
                const bool truncated = Wa + nw < seg_end;
                FJ_T(1)
                if (W >= seg_end) done = true;
                if (!done) {
                    __syncwarp();  // the previous chunk's readers are done with the shared arrays
                    if (lane == 0) {
                        const uint32_t bytes = ((nw + 3u) & ~3u) * 4u;  // (the slab's sections are padded: reading up to 3 rows past n is fine)
                        mbar_expect_tx(&sm.bar[0], 2u * bytes);
                        tma_load_1d(w_start, ix.s_start + Wa, bytes, &sm.bar[0], 0ull);
                        tma_load_1d(w_pme, ix.pme + Wa, bytes, &sm.bar[0], 0ull);
                        mbar_expect_tx(&sm.bar[1], 2u * bytes);
                        tma_load_1d(sm.w_end, ix.s_end + Wa, bytes, &sm.bar[1], 0ull);
                        tma_load_1d(sm.w_pos, ix.s_pos + Wa, bytes, &sm.bar[1], 0ull);
                    }
                    bool b_pending = true;  // the second copy has to be waited for on every path out of here
                    mbar_wait(&sm.bar[0], ph_a);
                    ph_a ^= 1u;
                    win_sentinels<FJ_WCAP>(w_start, nw, lane);
                    win_sentinels<FJ_WCAP>(w_pme, nw, lane);
                    __syncwarp();
                    FJ_T(2)
                    // ---- per row: i0 = #(prefix_max_end < qs), n_le = #(start <= qe), both relative to Wa --------------------
                    uint32_t i0[FJ_PV], cand[FJ_PV];
                    win_ranks<false, FJ_PV, FJ_WCAP>(w_pme, wlo, nw, qs, i0);
                    win_ranks<true, FJ_PV, FJ_WCAP>(w_start, wlo, nw, qe, cand);
                    FJ_T(3)
                    bool beyond = false;  // a row whose candidates may continue behind a truncated window
#pragma unroll
                    for (int j = 0; j < FJ_PV; j++) {
                        beyond = beyond | (cand[j] == nw);
                        cand[j] = cand[j] > i0[j] ? cand[j] - i0[j] : 0u;
                    }
                    uint32_t pj[FJ_PV];
                    const uint32_t n_slots = fj_row_prefix16(cand, lane, pj);
                    if (truncated && __any_sync(0xFFFFFFFFu, beyond)) fast = false;
                    else if (n_slots == 0) done = true;
                    else if (n_slots > (uint32_t)FJ_SLOTS) fast = false;
                    if (fast && !done) {
                        // row record {qs, i0 - prefix}: slot s of the row is window row s + (i0 - prefix)
#pragma unroll
                        for (int j = 0; j < FJ_PV; j++) sm.row2[fj_row(lane, j)] = make_uint2((uint32_t)qs[j], i0[j] - pj[j]);
                        fj_build_rowof(sm.rowof, n_slots, 0u, (uint32_t)FJ_ROWS, 0u, pj, cand, lane);  // (orders the stores above, too)
                        FJ_T(4)
                        mbar_wait(&sm.bar[1], ph_b);
                        ph_b ^= 1u;
                        b_pending = false;
                        // ---- test 32 candidate slots per step, compact, stage (over the searched arrays: every lane is past them) --
                        // (slots past n_slots read the zeroed tail of the map = row 0; their window row is clamped)
                        uint32_t run = 0;
                        for (uint32_t sb = 0; sb < n_slots; sb += 32 * FJ_U) {
                            uint32_t val[FJ_U];
                            bool mt[FJ_U];
#pragma unroll
                            for (int u = 0; u < FJ_U; u++) {
                                const uint32_t s = sb + u * 32 + lane;
                                const uint32_t r = sm.rowof[s];
                                const uint2 rw = sm.row2[r];
                                const uint32_t jrel = min(s + rw.y, (uint32_t)FJ_WCAP - 1u);
                                mt[u] = (s < n_slots) & (sm.w_end[jrel] >= (int32_t)rw.x);
                                val[u] = jrel | (r << 16);
                            }
#pragma unroll
                            for (int u = 0; u < FJ_U; u++) {
                                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, mt[u]);
                                const uint32_t o = run + __popc(bal & lt_mask);
                                if (mt[u] && o < (uint32_t)FJ_STAGE) sm.u.stage[o] = val[u];
                                run += __popc(bal);
                            }
                        }
                        FJ_T(5)
                        if (run > (uint32_t)FJ_STAGE) fast = false;  // more pairs than the staging area holds: count + direct emission
                        else { n_staged = run; total = run; done = true; FJ_N(9) }
                    }
                    if (b_pending) {  // (no candidates, or off to the slow path: the copy must have landed before the buffers are reused)
                        mbar_wait(&sm.bar[1], ph_b);
                        ph_b ^= 1u;
                    }
                }
            }
        }
        if (done) {
            // ---- the chunk's output slots, then the staged pairs as 128-bit stores ---------------------------------------------
            base = chunk_slots(sc, chunk, total, lane);
            FJ_T(6)
            if (n_staged) {
                __syncwarp();
                const uint32_t* stage = sm.u.stage;
                uint32_t* gb = build_idx + base;
                uint32_t* gp = probe_idx + base;
                const uint32_t r0 = (uint32_t)row0;
                const uint64_t room = capacity > base ? capacity - base : 0ull;
                const uint32_t n_out = (uint32_t)min((uint64_t)n_staged, room);
#if BRG_FJ_VEC
                // head up to the first 16-byte boundary of the output, 128-bit body, tail (vec_out: both columns are 16-byte aligned)
                const uint32_t head = vec_out ? min(n_out, (uint32_t)((4u - (uint32_t)(base & 3u)) & 3u)) : n_out;
                for (uint32_t k = lane; k < head; k += 32) {
                    const uint32_t v = stage[k];
                    gb[k] = sm.w_pos[v & 0xFFFFu];
                    gp[k] = r0 + (v >> 16);
                }
                const uint32_t n_vec = (n_out - head) >> 2;
                for (uint32_t x = lane; x < n_vec; x += 32) {
                    const uint32_t k = head + 4u * x;
                    const uint32_t v0 = stage[k], v1 = stage[k + 1], v2 = stage[k + 2], v3 = stage[k + 3];
                    const uint4 pb = make_uint4(sm.w_pos[v0 & 0xFFFFu], sm.w_pos[v1 & 0xFFFFu], sm.w_pos[v2 & 0xFFFFu], sm.w_pos[v3 & 0xFFFFu]);
                    const uint4 pp = make_uint4(r0 + (v0 >> 16), r0 + (v1 >> 16), r0 + (v2 >> 16), r0 + (v3 >> 16));
                    __stcs(reinterpret_cast<uint4*>(gb + k), pb);
                    __stcs(reinterpret_cast<uint4*>(gp + k), pp);
                }
                for (uint32_t k = head + 4u * n_vec + lane; k < n_out; k += 32) {
                    const uint32_t v = stage[k];
                    gb[k] = sm.w_pos[v & 0xFFFFu];
                    gp[k] = r0 + (v >> 16);
                }
#else
                // one 32-bit element per lane: every store instruction of the warp writes 128 contiguous bytes, and the staged words
                // are read without bank conflicts (four consecutive words per lane, as 128-bit stores need them, conflict 4-way; the
                // kernel is bound by the shared-memory pipe)
#pragma unroll 4
                for (uint32_t k = lane; k < n_out; k += 32) {
                    const uint32_t v = stage[k];
                    __stcs(gb + k, sm.w_pos[v & 0xFFFFu]);
                    __stcs(gp + k, r0 + (v >> 16));
                }
#endif
            }
            FJ_T(7)
        } else {
            // ---- slow path: per-row directory searches, count, output slots, emit pass ----------------------------------------
            uint32_t n_le[FJ_PV], i0[FJ_PV], cand[FJ_PV];
            bool act[FJ_PV];
            int32_t qm[FJ_PV];
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) act[j] = kk[j] < ix.n_keys;
            fj_dir_ranks8(ix.d_sstart, kk, qe, act, n_le);  // #(start <= qe)
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) {
                // first row with prefix_max_end >= qs = rank_le(qs - 1); qs == INT32_MIN: the segment's first row (n_le's floor)
                qm[j] = wsub(qs[j], 1);
                if (qs[j] == INT32_MIN) { act[j] = false; }
            }
            fj_dir_ranks8(ix.d_pme, kk, qm, act, i0);
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) {
                const bool valid = kk[j] < ix.n_keys;
                if (valid && qs[j] == INT32_MIN) i0[j] = load_keydir(ix.d_pme.kd, kk[j]).seg_begin;
                cand[j] = (valid && n_le[j] > i0[j]) ? n_le[j] - i0[j] : 0u;
            }
            // exclusive prefix of the candidate counts in row order (row = lane + 32 j), 64-bit: one warp scan per group of 32 rows
            uint64_t pj64[FJ_PV], all = 0;
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) {
                uint64_t incl = cand[j];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= (uint32_t)o) incl += y;
                }
                pj64[j] = all + incl - cand[j];
                all += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            const bool huge = all > 0xFFFFFFFFull;  // the chunk's candidate prefix does not fit 32 bits: rows are swept one by one
            uint32_t pj[FJ_PV];
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) pj[j] = (uint32_t)pj64[j];
            __syncwarp();  // the previous users of the shared arrays are done
#pragma unroll
            for (int j = 0; j < FJ_PV; j++) sm.u.row4[fj_row(lane, j)] = make_uint4(pj[j], i0[j], (uint32_t)qs[j], cand[j]);
            if (lane == 31) sm.u.row4[FJ_ROWS] = make_uint4((uint32_t)all, 0u, 0u, 0u);
            __syncwarp();
            if (all) {
                // The chunk's pair count (in ordered mode every later chunk waits for it).  Count identity (no build row with
                // start > end, qs <= qe): matches = #(start <= qe) - #(end < qs), one more directory search instead of a pass over
                // the candidate slots.
                bool plain = (ix.flags & BRG_IDX_INVERTED) == 0;
#pragma unroll
                for (int j = 0; j < FJ_PV; j++) plain = plain & !(qe[j] < qs[j] && cand[j] != 0u);
                if (__all_sync(0xFFFFFFFFu, plain)) {
                    uint32_t ended[FJ_PV];
#pragma unroll
                    for (int j = 0; j < FJ_PV; j++) act[j] = cand[j] != 0u && qs[j] != INT32_MIN;
                    fj_dir_ranks8(ix.d_eend, kk, qm, act, ended);  // #(end < qs) = #(end <= qs - 1), as a global index
                    uint64_t cnt = 0;
#pragma unroll
                    for (int j = 0; j < FJ_PV; j++) {
                        if (cand[j]) {
                            // (segments of e_end and s_start coincide: both are the key's rows)
                            const uint32_t e0 = act[j] ? ended[j] : load_keydir(ix.d_eend.kd, kk[j]).seg_begin;
                            cnt += n_le[j] - e0;
                        }
                    }
                    total = warp_sum_u64(cnt);
                } else {
                    total = fj_slow_pass<false>(ix, sm, sm.u.row4, pj, cand, huge, (uint32_t)row0, 0ull, 0ull, nullptr, nullptr, lane);
                }
            }
            base = chunk_slots(sc, chunk, total, lane);
            if (total) fj_slow_pass<true>(ix, sm, sm.u.row4, pj, cand, huge, (uint32_t)row0, base, capacity, build_idx, probe_idx, lane);
            __syncwarp();
            FJ_T(8) FJ_N(10)
        }
        if (lane == 0) {
            if (chunk_slots_out) { chunk_slots_out[2ull * chunk] = base; chunk_slots_out[2ull * chunk + 1] = total; }
            if (sc.ordered && chunk == n_chunks - 1) *total_out = base + total;
        }
        chunk++;
    }
    FJ_PROF_FLUSH
}

uint32_t pairs_join_chunk_rows() { return (uint32_t)FJ_ROWS; }
static inline uint64_t join_chunks(uint64_t n) { return (n + FJ_ROWS - 1) / FJ_ROWS; }
uint64_t pairs_join_state_bytes(uint64_t n) {  // the ticket + the cursor (2 words), one word per chunk, one per group of 32 chunks
    const uint64_t c = join_chunks(n);
    return (2 + c + (c + 31) / 32) * sizeof(uint64_t);
}

static inline bool aligned16j(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int launch_pairs_join(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n, uint64_t capacity,
                      int ordered, uint32_t* build_idx, uint32_t* probe_idx, uint64_t* d_total, uint64_t* chunk_slots_out, void* d_state,
                      cudaStream_t st) {
    const uint64_t n_chunks = join_chunks(n);
    if (n_chunks == 0) {
        BRG_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint64_t), st));
        return 0;
    }
    // [0]: the chunk ticket, [1]: the output cursor, [2 ..]: one look-back word per chunk, then one per group
    BRG_CUDA(cudaMemsetAsync(d_state, 0, ordered ? pairs_join_state_bytes(n) : 2 * sizeof(uint64_t), st));
    BRG_CUDA(cudaFuncSetAttribute((const void*)pairs_join_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(JoinWarp) * FJ_WARPS)));
    const uint64_t ctas = (n_chunks + FJ_WARPS - 1) / FJ_WARPS;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(ctas, (uint64_t)kSMs * BRG_FJ_CTAS));
    const int vec_in = aligned16j(s) && aligned16j(e) && (!key || aligned16j(key));
    const int vec_out = aligned16j(build_idx) && aligned16j(probe_idx);
    uint64_t* words = reinterpret_cast<uint64_t*>(d_state);
    const ScanState sc{words + 1, words + 2, words + 2 + n_chunks, (uint32_t)n_chunks, ordered};
    pairs_join_kernel<<<grid, FJ_WARPS * 32, sizeof(JoinWarp) * FJ_WARPS, st>>>(ix, key, s, e, n, reinterpret_cast<uint32_t*>(words), sc,
                                                                                capacity, build_idx, probe_idx, d_total, chunk_slots_out,
                                                                                vec_in, vec_out);
    BRG_LAUNCH_CHECK();
    if (!ordered) BRG_CUDA(cudaMemcpyAsync(d_total, words + 1, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));  // the cursor = the total
    return 0;
}

#ifdef BRG_FJ_PROF
// out[0..11]: clock sums of  0 ticket + probe loads + order checks, 1 window bounds, 2 window staging, 3 rank searches, 4 prefix +
// records + slot map, 5 slot tests, 6 look-back, 7 copy-out, 8 slow path (whole), 9 / 10 chunks on the fast / slow path
extern "C" __attribute__((visibility("default"))) int brg_debug_join_prof(unsigned long long* out, int reset) {
    if (out && cudaMemcpyFromSymbol(out, g_fj_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(g_fj_prof, z, sizeof(z)) != cudaSuccess) return 1;
    }
    return 0;
}
#endif

}  // namespace brg
