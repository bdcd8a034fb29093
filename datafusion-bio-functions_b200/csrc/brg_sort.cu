// brg_sort.cu -- K1: stable LSD radix sort of build rows for sm_100a.
//
// Replaces the reference's CPU sorts of the build side: `starts.sort_unstable(); ends.sort_unstable()`
// (interval_tree.rs:35-36), `by_start.sort_by(first,last,meta)` (nearest_index.rs:50-55), `by_end`
// (nearest_index.rs:77-82) and the sort inside COITree::new (interval_join.rs:760).
//
// Rows are a struct-of-arrays of up to 4 u32 words (key id, start, end, row id).  One pass = three launches:
//   (a) per-tile 256-bin digit histogram (warp-aggregated with match.any, shared-memory atomics),
//   (b) exclusive scan of the digit-major (digit, tile) table  -> global base of every (digit, tile) run,
//   (c) scatter: every tile ranks its rows stably per digit with warp-level match.any multisplit + per-warp
//       shared counters, reorders each word array through shared memory so that rows of one digit leave the
//       SM as contiguous runs, and writes them to their final place.
// HBM-bound: per pass each word array is read once and written once (plus one read of the digit array in (a)).
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

__global__ void __launch_bounds__(RS_THREADS, BRG_RS_MINB) radix_scatter_kernel(SortArrays arr, uint64_t n, SortPass pass,
                                                                   uint32_t n_tiles,
                                                                   const uint32_t* __restrict__ hist_scanned) {
    __shared__ uint32_t wc[RS_WARPS][256];  // per-warp digit counters -> per-warp exclusive offsets
    __shared__ uint32_t tile_start[256];    // exclusive prefix of the tile's digit counts
    __shared__ uint32_t gbase[256];         // global base of the (digit, tile) run minus tile_start
    __shared__ uint32_t s_dest[RS_TILE];
    __shared__ uint32_t s_val[RS_TILE];
    __shared__ uint32_t warp_tot[RS_WARPS];

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const uint64_t tile_base = (uint64_t)blockIdx.x * RS_TILE;
    const uint64_t warp_base = tile_base + (uint64_t)w * (RS_ITEMS * 32);
    const uint32_t n_tile_valid = (uint32_t)min((uint64_t)RS_TILE, n - tile_base);
    const uint32_t lanemask_lt = (1u << lane) - 1u;

    for (int i = t; i < RS_WARPS * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
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
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        uint32_t d = dig[j];
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
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
    gbase[t] = hist_scanned[(uint64_t)t * n_tiles + blockIdx.x] - excl;
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
            uint32_t v = (a == pass.src) ? keyv[j] : ((i < n) ? __ldg(in + i) : 0u);
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

uint64_t radix_hist_words(uint64_t n) {
    uint64_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
    return 256ull * (n_tiles ? n_tiles : 1);
}

int radix_sort_rows(uint32_t* const* a, uint32_t* const* b, int n_arr, uint64_t n, const SortPass* passes,
                    int n_passes, uint32_t* d_hist, uint64_t hist_words, uint32_t* d_scan_scratch, cudaStream_t st,
                    int* res) {
    *res = 0;
    if (n == 0 || n_passes == 0) return 0;
    if (n_arr > RS_MAX_ARR) { set_error("radix_sort_rows: too many arrays"); return 1; }
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
        radix_scatter_kernel<<<n_tiles, RS_THREADS, 0, st>>>(arr, n, passes[p], n_tiles, d_hist);
        BRG_LAUNCH_CHECK();
        cur ^= 1;
    }
    *res = cur;
    return 0;
}

}  // namespace brg
