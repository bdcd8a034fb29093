// brg_sort.cu -- K1: stable LSD radix sort of build rows for sm_100a.
//
// Replaces the reference's CPU sorts of the build side: `starts.sort_unstable(); ends.sort_unstable()`
// (interval_tree.rs:35-36), `by_start.sort_by(first,last,meta)` (nearest_index.rs:50-55), `by_end`
// (nearest_index.rs:77-82) and the sort inside COITree::new (interval_join.rs:760).
//
// Rows are a struct-of-arrays of up to 4 u32 words (key id, start, end, row id).  Single-pass ("onesweep") scheme:
//   (a) ONE kernel reads the digit source arrays once and builds the 256-bin digit histogram of EVERY pass,
//   (b) one launch per pass: tiles are taken in ticket order; every tile ranks its rows stably per digit with
//       warp-level match.any multisplit + per-warp shared counters, publishes its 256 digit counts, and obtains the
//       number of rows with the same digit in earlier tiles by a decoupled look-back over the tiles in front of it
//       (thread d follows digit d) -- no per-tile histogram pass, no (digit x tile) table scan; then it reorders each
//       word array through shared memory so that rows of one digit leave the SM as contiguous runs.
// Per pass each word array is read once and written once.  (Round 1 ran three launches per pass -- per-tile histogram,
// table scan, scatter: 0.29 ms per pass and 10M rows against 0.1 ms now.)  Inputs of 2^30 rows or more keep that scheme
// (the look-back packs count and state into 32 bits).
#include <algorithm>
#include <cstdlib>

#include "brg_internal.cuh"

namespace brg {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef BRG_RS_ITEMS
#define BRG_RS_ITEMS 16
#endif
#ifndef BRG_RS_MINB
#define BRG_RS_MINB 3  // resident CTAs per SM of the scatter kernel: 80 registers; it is latency-bound (ncu: issue slots 9 % busy at 2 CTAs)
#endif
#ifndef BRG_RS_PREFETCH
#define BRG_RS_PREFETCH 1
#endif
constexpr int RS_ITEMS = BRG_RS_ITEMS;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 rows per CTA
constexpr int RS_MAX_ARR = 4;

struct SortArrays {
    const uint32_t* in[RS_MAX_ARR];
    uint32_t* out[RS_MAX_ARR];
    int n_arr;
};

__device__ __forceinline__ uint32_t digit_of(uint32_t v, const SortPass& p) {
    if (v > p.clamp) v = p.clamp;
    if (p.flip_sign) v ^= 0x80000000u;
    return (v >> p.shift) & 0xFFu;
}

__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint32_t* __restrict__ src, uint64_t n,
                                                                SortPass pass, uint32_t n_tiles,
                                                                uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    const int t = threadIdx.x;
    h[t] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
    // all loads first (the kernel was bound by the latency of load -> match.any -> atomic chains, ncu: issue slots 11 % busy)
    uint32_t d[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const uint64_t i = base + (uint64_t)j * RS_THREADS + t;
        d[j] = i < n ? digit_of(__ldg(src + i), pass) : 0x100u;
    }
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        // a warp whose 32 digits agree (the high bytes of sorted or clustered coordinates) adds once; otherwise one shared
        // atomic per lane: random digits rarely collide, and match.any costs more than the few serialised collisions
        int all_same;
        __match_all_sync(0xFFFFFFFFu, d[j], &all_same);
        if (all_same) {
            if ((t & 31) == 0 && d[j] < 256u) atomicAdd(&h[d[j]], 32u);
        } else if (d[j] < 256u) {
            atomicAdd(&h[d[j]], 1u);
        }
    }
    __syncthreads();
    hist[(uint64_t)t * n_tiles + blockIdx.x] = h[t];  // digit-major table
}

// N_ARR == 2 (key + one value array, the sorts of a count-only build): both arrays are loaded into registers up front, so the
// payload loads are in flight while the ranks are computed (the generic version, N_ARR = 0, loads each payload array between
// two barriers and exposes its latency: ncu 24 % warps active, issue slots 9 % busy).
constexpr uint32_t OS_AGG = 1u << 30, OS_PRE = 2u << 30, OS_VAL = OS_AGG - 1u;  // look-back state of one (tile, digit)
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// hist_scanned != nullptr: legacy scheme (tile = blockIdx.x, bases from the scanned (digit, tile) table).
// Otherwise onesweep: ghist = this pass's 256 global digit counts, status = n_tiles x 256 zeroed words, ticket = zeroed counter.
template <int N_ARR>
__global__ void __launch_bounds__(RS_THREADS, BRG_RS_MINB) radix_scatter_kernel(SortArrays arr, uint64_t n, SortPass pass,
                                                                   uint32_t n_tiles,
                                                                   const uint32_t* __restrict__ hist_scanned,
                                                                   const uint32_t* __restrict__ ghist, uint32_t* status,
                                                                   uint32_t* ticket) {
    __shared__ uint32_t wc[RS_WARPS][256];  // per-warp digit counters -> per-warp exclusive offsets
    __shared__ uint32_t tile_start[256];    // exclusive prefix of the tile's digit counts
    __shared__ uint32_t gbase[256];         // global base of the (digit, tile) run minus tile_start
    __shared__ uint32_t s_dest[RS_TILE];
    __shared__ uint32_t s_val[RS_TILE];
    __shared__ uint32_t warp_tot[RS_WARPS];

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    __shared__ uint32_t s_tile;
    if (hist_scanned == nullptr) {  // tiles in ticket order: every tile in front of this one is running or done
        if (t == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
    }
    const uint32_t tile = hist_scanned ? blockIdx.x : s_tile;
    const uint64_t tile_base = (uint64_t)tile * RS_TILE;
    const uint64_t warp_base = tile_base + (uint64_t)w * (RS_ITEMS * 32);
    const uint32_t n_tile_valid = (uint32_t)min((uint64_t)RS_TILE, n - tile_base);
    const uint32_t lanemask_lt = (1u << lane) - 1u;

    for (int i = t; i < RS_WARPS * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
#if BRG_RS_PREFETCH
    if (N_ARR != 2) {
        // The generic version loads each payload array only when it drains it, between two barriers: start those lines towards L2
        // now (16 KB per array and tile = 128 lines; no registers to hold them in).
        if (t < 128) {
            for (int a = 0; a < arr.n_arr; a++) {
                const uint64_t i = tile_base + (uint64_t)t * 32;
                if (a != pass.src && i < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(arr.in[a] + i));
            }
        }
    }
#endif
    __syncthreads();

    // tile order = index order: warp w owns rows [w*512, (w+1)*512), round j covers 32 consecutive rows
    uint32_t dig[RS_ITEMS], rank[RS_ITEMS], keyv[RS_ITEMS];
    const uint32_t* src = arr.in[pass.src];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        uint64_t i = warp_base + (uint64_t)j * 32 + lane;
        keyv[j] = (i < n) ? __ldg(src + i) : 0u;
        dig[j] = (i < n) ? digit_of(keyv[j], pass) : 0xFFu;  // padding sorts after every valid row of the tile
    }
    uint32_t pay[RS_ITEMS];  // N_ARR == 2: the one array that is not the digit source (pass.src is 0 or 1)
    if (N_ARR == 2) {
        const uint32_t* in = arr.in[1 - pass.src];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const uint64_t i = warp_base + (uint64_t)j * 32 + lane;
            pay[j] = (i < n) ? __ldg(in + i) : 0u;
        }
    }
    // (all matches first: they are independent of the counter updates below, which form one dependent chain through shared memory)
    uint32_t peer_mask[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) peer_mask[j] = __match_any_sync(0xFFFFFFFFu, dig[j]);
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        uint32_t d = dig[j];
        uint32_t peers = peer_mask[j];
        int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) {
            base = wc[w][d];
            wc[w][d] = base + __popc(peers);
        }
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        rank[j] = base + __popc(peers & lanemask_lt);
        __syncwarp();
    }
    __syncthreads();

    // per digit: exclusive prefix over warps, tile total
    uint32_t total;
    {
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            uint32_t c = wc[ww][t];
            wc[ww][t] = run;
            run += c;
        }
        total = run;
    }
    // exclusive scan of the 256 digit totals (warp shuffle scan + warp totals)
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ww++)
        if (ww < w) woff += warp_tot[ww];
    uint32_t excl = woff + incl - total;
    tile_start[t] = excl;
    if (hist_scanned) {
        gbase[t] = hist_scanned[(uint64_t)t * n_tiles + tile] - excl;
    } else {
        // first row of digit t in the output = rows with a smaller digit (scan of the global histogram) ...
        const uint32_t g = __ldg(ghist + t);
        uint32_t gi = g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, gi, o);
            if (lane >= o) gi += v;
        }
        __syncthreads();  // (warp_tot was read above by every thread)
        if (lane == 31) warp_tot[w] = gi;
        __syncthreads();
        uint32_t goff = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++)
            if (ww < w) goff += warp_tot[ww];
        const uint32_t dbase = goff + gi - g;
        // ... + rows with digit t in the tiles in front of this one: decoupled look-back, thread t follows digit t
        uint32_t prev = 0;
        uint32_t* mine = status + (uint64_t)tile * 256u + t;
        if (tile == 0) {
            atomicExch(mine, OS_PRE | total);
        } else {
            atomicExch(mine, OS_AGG | total);
            // four tiles per round trip: the loads of one round are independent (tiles start in lock step, so the nearest
            // inclusive prefix lies a dozen tiles back; walking them one L2 latency at a time was a fifth of the kernel)
            bool open = true;
            for (int64_t j = (int64_t)tile - 1; open && j >= 0; j -= 4) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = (j - u >= 0) ? ld_volatile_u32(status + (uint64_t)(j - u) * 256u + t) : OS_PRE;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (!open) break;
                    while ((v[u] & (OS_AGG | OS_PRE)) == 0u) v[u] = ld_volatile_u32(status + (uint64_t)(j - u) * 256u + t);
                    prev += v[u] & OS_VAL;
                    if (v[u] & OS_PRE) open = false;
                }
            }
            atomicExch(mine, OS_PRE | (prev + total));
        }
        gbase[t] = dbase + prev - excl;
    }
    __syncthreads();

    // tile-local sorted position of every row and its global destination
    uint32_t pos[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        uint32_t d = dig[j];
        uint32_t p = tile_start[d] + wc[w][d] + rank[j];
        pos[j] = p;
        s_dest[p] = gbase[d] + p;
    }
    // one word array at a time: registers -> shared (sorted order) -> global (runs of one digit are contiguous)
    for (int a = 0; a < arr.n_arr; a++) {
        __syncthreads();  // s_dest complete / previous drain finished
        const uint32_t* in = arr.in[a];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            uint64_t i = warp_base + (uint64_t)j * 32 + lane;
            uint32_t v;
            if (a == pass.src) v = keyv[j];
            else if (N_ARR == 2) v = pay[j];
            else v = (i < n) ? __ldg(in + i) : 0u;
            s_val[pos[j]] = v;
        }
        __syncthreads();
        uint32_t* out = arr.out[a];
#pragma unroll
        for (int k = 0; k < RS_ITEMS; k++) {
            uint32_t q = k * RS_THREADS + t;
            if (q < n_tile_valid) out[s_dest[q]] = s_val[q];
        }
    }
}

// (a) of the onesweep scheme: the digit histograms of ALL passes in one read of the source arrays.
constexpr int OS_MAX_PASSES = 12;
struct HistPlan {
    const uint32_t* src[OS_MAX_PASSES];
    SortPass pass[OS_MAX_PASSES];
    int n_passes;
};
__global__ void __launch_bounds__(RS_THREADS) radix_global_hist_kernel(HistPlan plan, uint64_t n, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t h[OS_MAX_PASSES][256];
    const int t = threadIdx.x;
    for (int p = 0; p < plan.n_passes; p++) h[p][t] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * RS_THREADS;
    for (uint64_t i0 = (uint64_t)blockIdx.x * RS_THREADS; i0 < n; i0 += stride * 4) {
        uint64_t idx[4];
#pragma unroll
        for (int u = 0; u < 4; u++) idx[u] = i0 + (uint64_t)u * stride + t;
        const uint32_t* last_src = nullptr;
        uint32_t v[4] = {0, 0, 0, 0};
        for (int p = 0; p < plan.n_passes; p++) {
            if (plan.src[p] != last_src) {  // consecutive passes over the same array share the loads
                last_src = plan.src[p];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = idx[u] < n ? __ldg(last_src + idx[u]) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t d = idx[u] < n ? digit_of(v[u], plan.pass[p]) : 0x100u;
                // a warp whose 32 digits agree (the high bytes of sorted or clustered coordinates) adds once
                int all_same;
                __match_all_sync(0xFFFFFFFFu, d, &all_same);
                if (all_same) {
                    if ((t & 31) == 0 && d < 256u) atomicAdd(&h[p][d], 32u);
                } else if (d < 256u) {
                    atomicAdd(&h[p][d], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int p = 0; p < plan.n_passes; p++)
        if (h[p][t]) atomicAdd(ghist + p * 256 + t, h[p][t]);
}

uint64_t radix_hist_words(uint64_t n) {
    // onesweep: per pass a (tile x 256) look-back table, plus the global histograms and the tile tickets
    uint64_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
    return (uint64_t)OS_MAX_PASSES * (256ull * (n_tiles ? n_tiles : 1) + 256ull + 1ull) + 64ull;
}

static int radix_sort_rows_legacy(uint32_t* const* a, uint32_t* const* b, int n_arr, uint64_t n, const SortPass* passes,
                                  int n_passes, uint32_t* d_hist, uint64_t hist_words, uint32_t* d_scan_scratch, cudaStream_t st,
                                  int* res);

int radix_sort_rows(uint32_t* const* a, uint32_t* const* b, int n_arr, uint64_t n, const SortPass* passes,
                    int n_passes, uint32_t* d_hist, uint64_t hist_words, uint32_t* d_scan_scratch, cudaStream_t st,
                    int* res) {
    *res = 0;
    if (n == 0 || n_passes == 0) return 0;
    if (n_arr > RS_MAX_ARR) { set_error("radix_sort_rows: too many arrays"); return 1; }
    static const bool legacy = getenv("BRG_SORT_LEGACY") != nullptr;  // the round-1 scheme (measured alternative)
    if (legacy || n >= (1ull << 30) || n_passes > OS_MAX_PASSES)
        return radix_sort_rows_legacy(a, b, n_arr, n, passes, n_passes, d_hist, hist_words, d_scan_scratch, st, res);
    const uint32_t n_tiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    const uint64_t per_pass = 256ull * n_tiles;
    if (hist_words < radix_hist_words(n)) { set_error("radix_sort_rows: histogram scratch too small"); return 1; }
    // scratch layout: [status: n_passes x n_tiles x 256][ghist: n_passes x 256][tickets: n_passes]
    uint32_t* status = d_hist;
    uint32_t* ghist = d_hist + (uint64_t)n_passes * per_pass;
    uint32_t* tickets = ghist + (uint64_t)n_passes * 256;
    BRG_CUDA(cudaMemsetAsync(d_hist, 0, ((uint64_t)n_passes * (per_pass + 256 + 1)) * 4, st));
    HistPlan plan;
    plan.n_passes = n_passes;
    for (int p = 0; p < n_passes; p++) {
        plan.pass[p] = passes[p];
        plan.src[p] = a[passes[p].src];  // the histogram of a digit does not depend on the row order: read the input set
    }
    {
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + RS_THREADS * 4 - 1) / (RS_THREADS * 4), (uint64_t)kSMs * 8));
        radix_global_hist_kernel<<<grid, RS_THREADS, 0, st>>>(plan, n, ghist);
        BRG_LAUNCH_CHECK();
    }
    int cur = 0;
    for (int p = 0; p < n_passes; p++) {
        SortArrays arr;
        arr.n_arr = n_arr;
        for (int i = 0; i < RS_MAX_ARR; i++) {
            arr.in[i] = i < n_arr ? (cur == 0 ? a[i] : b[i]) : nullptr;
            arr.out[i] = i < n_arr ? (cur == 0 ? b[i] : a[i]) : nullptr;
        }
        if (n_arr == 2)
            radix_scatter_kernel<2><<<n_tiles, RS_THREADS, 0, st>>>(arr, n, passes[p], n_tiles, nullptr, ghist + p * 256, status + p * per_pass, tickets + p);
        else
            radix_scatter_kernel<0><<<n_tiles, RS_THREADS, 0, st>>>(arr, n, passes[p], n_tiles, nullptr, ghist + p * 256, status + p * per_pass, tickets + p);
        BRG_LAUNCH_CHECK();
        cur ^= 1;
    }
    *res = cur;
    return 0;
}

static int radix_sort_rows_legacy(uint32_t* const* a, uint32_t* const* b, int n_arr, uint64_t n, const SortPass* passes,
                                  int n_passes, uint32_t* d_hist, uint64_t hist_words, uint32_t* d_scan_scratch, cudaStream_t st,
                                  int* res) {
    *res = 0;
    const uint32_t n_tiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    if (hist_words < 256ull * n_tiles) { set_error("radix_sort_rows: histogram scratch too small"); return 1; }
    int cur = 0;
    for (int p = 0; p < n_passes; p++) {
        SortArrays arr;
        arr.n_arr = n_arr;
        for (int i = 0; i < RS_MAX_ARR; i++) {
            arr.in[i] = i < n_arr ? (cur == 0 ? a[i] : b[i]) : nullptr;
            arr.out[i] = i < n_arr ? (cur == 0 ? b[i] : a[i]) : nullptr;
        }
        radix_hist_kernel<<<n_tiles, RS_THREADS, 0, st>>>(arr.in[passes[p].src], n, passes[p], n_tiles, d_hist);
        BRG_LAUNCH_CHECK();
        int rc = exclusive_scan_u32(d_hist, d_hist, 256ull * n_tiles, d_scan_scratch, st);
        if (rc) return rc;
        if (n_arr == 2) radix_scatter_kernel<2><<<n_tiles, RS_THREADS, 0, st>>>(arr, n, passes[p], n_tiles, d_hist, nullptr, nullptr, nullptr);
        else radix_scatter_kernel<0><<<n_tiles, RS_THREADS, 0, st>>>(arr, n, passes[p], n_tiles, d_hist, nullptr, nullptr, nullptr);
        BRG_LAUNCH_CHECK();
        cur ^= 1;
    }
    *res = cur;
    return 0;
}

}  // namespace brg
