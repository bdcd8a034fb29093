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
    if (ix.nb.rec) {  // nearest / pair bins: both ranks from one 32-byte record (see IndexView::nb)
        const KeyDir k = load_keydir(ix.nb.kd, c);
        uint32_t n_le, p0;
        const uint32_t missing = bins_pair_ranks(ix.nb, k, qe, qs, &n_le, &p0);
        if (missing & 1u) n_le = rank_le_kd(ix.d_sstart, load_keydir(ix.d_sstart.kd, c), qe);
        if (missing & 2u) p0 = rank_lt_kd(ix.d_pme, load_keydir(ix.d_pme.kd, c), qs);
        if (p0 < n_le) { *i0 = p0; *i1 = n_le; }
        return;
    }
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
//   pass 1  pairs_ranges_kernel_tiled : per row, the candidate range [i0, n_le) of the start-sorted build rows
//           (i0 = first row whose prefix_max_end >= qs, n_le = #(start <= qe)), written as one u64 per row; per chunk,
//           the number of pairs = sum of  #(start <= qe) - #(end < qs)  (the count_overlaps identity, exact for build
//           rows with start <= end and probes with qs <= qe; other rows count by walking their range).
//           Probe columns arrive through the per-warp TMA rings of brg_tile.cuh, all ranks of a chunk come from
//           inverse searches keyed on the (sorted) probe starts.  Chunks that do not fit the pattern go to a
//           one-row-per-thread kernel through the to-do list.
//   scan    exclusive scan of the n/256 chunk totals (tiny)
//   pass 2  pairs_emit_kernel_flat : no searching.  A warp FLATTENS the candidate slots of its chunk (row-major, so the
//           output stays probe-major): a byte map slot -> row is built with one scattered store per row and a packed
//           prefix-max; then 32 slots are tested per step (`end >= qs`), compacted with a ballot and written as one
//           contiguous run of build_idx / probe_idx.  No lane ever walks a range alone, so skewed rows cost nothing
//           extra and the stores of a step are adjacent (the per-row emit kernel touched every 32 B sector up to eight
//           times).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t walk_count(const IndexView& ix, uint32_t i0, uint32_t i1, int32_t qs) {
    uint32_t c = 0;
    for (uint32_t i = i0; i < i1; i++) c += (__ldg(ix.s_end + i) >= qs) ? 1u : 0u;
    return c;
}

#ifndef BRG_PR_CTAS
#define BRG_PR_CTAS 2
#endif
__global__ void __launch_bounds__(TP_THREADS, BRG_PR_CTAS) pairs_ranges_kernel_tiled(IndexView ix, const uint32_t* __restrict__ key,
                                                                           const int32_t* __restrict__ start,
                                                                           const int32_t* __restrict__ end, uint64_t n_rows,
                                                                           uint64_t* __restrict__ row_rng,
                                                                           uint64_t* __restrict__ chunk_tot, TodoList todo) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ProbeRing* ring = reinterpret_cast<ProbeRing*>(smem_raw);
    uint32_t* aux0 = ring->w[threadIdx.x >> 5].aux[0];
    uint32_t* aux1 = ring->w[threadIdx.x >> 5].aux[1];
    uint32_t* aux2 = ring->w[threadIdx.x >> 5].aux[2];
    ProbePipe pipe(ring, key, start, end, n_rows);  // n_rows is a multiple of 256: every chunk is full
    ProbeTile t;
    WarpCursorFar cs, cp, ce;  // all three are served by single searches
    cs.reset(); cp.reset(); ce.reset();
    KeyDir ks, kp, ke;
    uint32_t kd_key = 0xFFFFFFFFu;
    uint32_t fails = 0;
    while (pipe.next(t)) {
        pad_tile(t);
        uint32_t kuni;
        bool ok = warp_uniform_key(t, ix.n_keys, &kuni);
        uint32_t n_le[PV], i0[PV], ended[PV];
        if (ok) {
            if (kuni != kd_key) {
                kd_key = kuni;
                ks = load_keydir(ix.d_sstart.kd, kuni);
                kp = load_keydir(ix.d_pme.kd, kuni);
                ke = load_keydir(ix.d_eend.kd, kuni);
            }
            const bool mono_s = chunk_mono(t.qs, t.full), mono_e = chunk_mono(t.qe, t.full);
            // the two start-keyed searches interleaved when both windows are short (as in the nearest kernel), else one by one
            const bool dual = mono_s && warp_window_ranks2<true, true>(cp, ix.d_pme, kp, t.qs, t.s_qs, aux1, i0, ce, ix.d_eend, ke, t.qs,
                                                                       t.s_qs, aux2, ended, kuni);
            ok = warp_ranks_le_end(cs, ix.d_sstart, ks, kuni, t.qs, t.qe, t.active, mono_s, mono_e, t.s_qs, t.s_qe, aux0, n_le);
            if (!dual) {
                ok = ok && warp_window_ranks<true>(cp, ix.d_pme, kp, kuni, t.qs, t.active, mono_s, t.s_qs, aux1, i0);
                ok = ok && warp_window_ranks<true>(ce, ix.d_eend, ke, kuni, t.qs, t.active, mono_s, t.s_qs, aux2, ended);
            }
        }
        if (ok) {
            // (a row matches < 2^32 build rows, but the 256 rows of a chunk together may not fit 32 bits -- a huge pile of duplicate
            //  build rows under wide probes: the halves of the counts are summed separately and recombined as u64)
            uint32_t tot_lo = 0, tot_hi = 0;
#pragma unroll
            for (int j = 0; j < PV; j++) {
                // start <= end on the build side (checked at build time): a row that ends before qs also starts before qe
                uint32_t c = n_le[j] - ended[j];
                if (t.qe[j] < t.qs[j]) c = (i0[j] < n_le[j]) ? walk_count(ix, i0[j], n_le[j], t.qs[j]) : 0u;  // unguarded predicate
                tot_lo += c & 0xFFFFu;
                tot_hi += c >> 16;
            }
            long long rng[PV];
#pragma unroll
            for (int j = 0; j < PV; j++) rng[j] = (long long)((uint64_t)i0[j] | ((uint64_t)n_le[j] << 32));
            store_chunk_i64(aux0, reinterpret_cast<int64_t*>(row_rng) + (t.row0 - (threadIdx.x & 31) * PV), rng);  // (aux0 + aux1)
            tot_lo = __reduce_add_sync(0xFFFFFFFFu, tot_lo);
            tot_hi = __reduce_add_sync(0xFFFFFFFFu, tot_hi);
            if ((threadIdx.x & 31) == 0) chunk_tot[t.chunk] = (uint64_t)tot_lo + ((uint64_t)tot_hi << 16);
            fails = 0;
        } else {
            cs.valid = false; cp.valid = false; ce.valid = false;
            todo_push(todo, t.chunk);
            fails++;
        }
        if (fails >= TP_BAIL || cs.wide || cp.wide || ce.wide) { pipe.bail_out(todo); break; }
        pipe.release();
    }
}

// deferred chunks (and the partial last chunk, and everything when the tiled kernel cannot run): one row per thread
__global__ void __launch_bounds__(TP_WTILE) pairs_ranges_todo_kernel(IndexView ix, const uint32_t* __restrict__ key,
                                                                     const int32_t* __restrict__ start,
                                                                     const int32_t* __restrict__ end, uint64_t n_rows,
                                                                     uint64_t* __restrict__ row_rng,
                                                                     uint64_t* __restrict__ chunk_tot, TodoList todo,
                                                                     uint32_t all_chunks, uint32_t tail_chunk) {
    __shared__ uint64_t wsum[TP_WTILE / 32];
    // the per-key constants of the three rank directories: every row needs all three, and a table has a few dozen keys
    constexpr uint32_t SK = 64;
    __shared__ KeyDir skd[3][SK];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_list = todo.count ? *todo.count : all_chunks;
    const uint32_t n_items = n_list + (tail_chunk != 0xFFFFFFFFu ? 1u : 0u);
    if (blockIdx.x >= n_items) return;
    const bool staged = ix.n_keys <= SK;
    if (staged) {
        const uint32_t words = ix.n_keys * (uint32_t)(sizeof(KeyDir) / 4);
        // (with the nearest / pair bins their constants take the first slot: the directories of s_start and pme are then read
        //  only behind a record that overflowed)
        const uint32_t* g[3] = {reinterpret_cast<const uint32_t*>(ix.nb.rec ? ix.nb.kd : ix.d_sstart.kd),
                                reinterpret_cast<const uint32_t*>(ix.d_pme.kd), reinterpret_cast<const uint32_t*>(ix.d_eend.kd)};
        for (int a = 0; a < 3; a++)
            for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) reinterpret_cast<uint32_t*>(skd[a])[i] = __ldg(g[a] + i);
        __syncthreads();
    }
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t chunk = item < n_list ? (todo.count ? todo.chunks[item] : item) : tail_chunk;
        const uint64_t i = (uint64_t)chunk * TP_WTILE + threadIdx.x;
        uint32_t cnt = 0;
        if (i < n_rows) {
            const int32_t qs = ld_stream(start + i), qe = ld_stream(end + i);
            const uint32_t c = key ? ld_stream(key + i) : 0u;
            uint32_t i0 = 0, i1 = 0;
            if (c < ix.n_keys && ix.nb.rec) {
                // nearest / pair bins: n_le and p0 from one 32-byte record; the count identity needs #(end < qs) on top
                const KeyDir kb = staged ? skd[0][c] : load_keydir(ix.nb.kd, c);
                uint32_t n_le, p0;
                const uint32_t missing = bins_pair_ranks(ix.nb, kb, qe, qs, &n_le, &p0);
                if (missing & 1u) n_le = rank_le_kd(ix.d_sstart, load_keydir(ix.d_sstart.kd, c), qe);
                if (missing & 2u) p0 = rank_lt_kd(ix.d_pme, staged ? skd[1][c] : load_keydir(ix.d_pme.kd, c), qs);
                if (p0 < n_le) {
                    i0 = p0; i1 = n_le;
                    // (the bins exist only without inverted build rows: the identity holds for every proper probe)
                    if (!(qe < qs)) cnt = n_le - rank_lt_kd(ix.d_eend, staged ? skd[2][c] : load_keydir(ix.d_eend.kd, c), qs);
                    else cnt = walk_count(ix, i0, i1, qs);
                }
            } else if (c < ix.n_keys) {
                // the three rank queries of the row are independent: directory steps together, bin searches in lock step
                //   n_le = #(start <= qe),  p0 = first row with prefix_max_end >= qs,  ended = #(end < qs)
                const KeyDir ks = staged ? skd[0][c] : load_keydir(ix.d_sstart.kd, c), kp = staged ? skd[1][c] : load_keydir(ix.d_pme.kd, c),
                             ke = staged ? skd[2][c] : load_keydir(ix.d_eend.kd, c);
                uint32_t n_le, h0, p0, h1, ended, h2;
                rank_le_bin(ix.d_sstart, ks, qe, &n_le, &h0);
                rank_lt_bin(ix.d_pme, kp, qs, &p0, &h1);
                rank_lt_bin(ix.d_eend, ke, qs, &ended, &h2);
                rank_search3(ix.d_sstart.A, qe, n_le, h0, ix.d_pme.A, wsub(qs, 1), p0, h1, ix.d_eend.A, wsub(qs, 1), ended, h2);
                if (p0 < n_le) {  // (rows before p0 end before qs, rows from n_le on start after qe)
                    i0 = p0; i1 = n_le;
                    // count identity (no build row with start > end, qs <= qe): matches = #(start <= qe) - #(end < qs);
                    // otherwise count by walking the candidate range
                    if (!(ix.flags & BRG_IDX_INVERTED) && !(qe < qs)) cnt = n_le - ended;
                    else cnt = walk_count(ix, i0, i1, qs);
                }
            }
            row_rng[i] = (uint64_t)i0 | ((uint64_t)i1 << 32);
        }
        const uint32_t c_lo = __reduce_add_sync(0xFFFFFFFFu, cnt & 0xFFFFu), c_hi = __reduce_add_sync(0xFFFFFFFFu, cnt >> 16);
        if (lane == 0) wsum[warp] = (uint64_t)c_lo + ((uint64_t)c_hi << 16);  // (u64: see pairs_ranges_kernel_tiled)
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t total = 0;
#pragma unroll
            for (int w = 0; w < TP_WTILE / 32; w++) total += wsum[w];
            chunk_tot[chunk] = total;
        }
        __syncthreads();
    }
}

constexpr int PF_THREADS = 256;
constexpr int PF_WARPS = PF_THREADS / 32;
#ifndef BRG_PF_PREFETCH
#define BRG_PF_PREFETCH 1
#endif
#ifndef BRG_PF_U
#define BRG_PF_U 4
#endif
#ifndef BRG_PF_CTAS
#define BRG_PF_CTAS 4
#endif
constexpr int PF_U = BRG_PF_U;  // slot steps in flight per iteration of the emit loop
constexpr int PF_CAP = 2304;  // candidate slots a warp flattens at a time (bytes of the slot -> row map); sized so that 4 CTAs fit an SM
// row r lives at slot r + r / 8: a lane writes the records of its 8 consecutive rows, and unpadded the 32 lanes of one such
// store instruction (128 bytes apart) would all hit the same banks -- 32 wavefronts per store, 250 conflict wavefronts per chunk
__device__ __forceinline__ uint32_t flat_row_slot(uint32_t r) { return r + (r >> 3); }
struct __align__(16) FlatWarp {
    uint4 row[TP_WTILE + TP_WTILE / 8 + 2];  // per row {P = exclusive prefix of the candidate counts, i0, qs, -}; row 256: .x = total
    uint8_t rowof[PF_CAP];
};

__global__ void __launch_bounds__(PF_THREADS, BRG_PF_CTAS) pairs_emit_kernel_flat(IndexView ix, const int32_t* __restrict__ start,
                                                                        uint64_t n_rows, uint64_t n_chunks,
                                                                        const uint64_t* __restrict__ row_rng,
                                                                        const uint64_t* __restrict__ chunk_base,
                                                                        uint32_t* __restrict__ build_idx,
                                                                        uint32_t* __restrict__ probe_idx, int vec_ok) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FlatWarp& sm = reinterpret_cast<FlatWarp*>(smem_raw)[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint64_t n_warps = (uint64_t)gridDim.x * PF_WARPS;
    for (uint64_t chunk = (uint64_t)blockIdx.x * PF_WARPS + (threadIdx.x >> 5); chunk < n_chunks; chunk += n_warps) {
        const uint64_t base = __ldg(chunk_base + chunk);
        if (__ldg(chunk_base + chunk + 1) == base) continue;  // no pairs in this chunk (warp-uniform)
        const uint64_t row = chunk * TP_WTILE + lane * PV;
        uint32_t i0[PV], cand[PV];
        int32_t qs[PV];
        if (vec_ok && row + PV <= n_rows) {  // (vec_ok: both columns are 16-byte aligned)
            const ulonglong2* pr = reinterpret_cast<const ulonglong2*>(row_rng + row);
            const int4* pq = reinterpret_cast<const int4*>(start + row);
#pragma unroll
            for (int j = 0; j < PV; j += 2) {
                const ulonglong2 v = __ldcs(pr + j / 2);
                i0[j] = (uint32_t)v.x; cand[j] = (uint32_t)(v.x >> 32);
                i0[j + 1] = (uint32_t)v.y; cand[j + 1] = (uint32_t)(v.y >> 32);
            }
            const int4 a = __ldcs(pq), b = __ldcs(pq + 1);
            qs[0] = a.x; qs[1] = a.y; qs[2] = a.z; qs[3] = a.w; qs[4] = b.x; qs[5] = b.y; qs[6] = b.z; qs[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < PV; j++) {
                const bool v = row + j < n_rows;
                const uint64_t r = v ? __ldcs(row_rng + row + j) : 0ull;
                i0[j] = (uint32_t)r; cand[j] = (uint32_t)(r >> 32);
                qs[j] = v ? __ldcs(start + row + j) : 0;
            }
        }
        uint32_t lane_tot = 0;
#pragma unroll
        for (int j = 0; j < PV; j++) { cand[j] = cand[j] > i0[j] ? cand[j] - i0[j] : 0u; lane_tot += cand[j]; }
#if BRG_PF_PREFETCH
        {   // the build rows the chunk is going to gather: prefetched now, needed after the slot map is built
            uint32_t wlo = 0xFFFFFFFFu, whi = 0u;
#pragma unroll
            for (int j = 0; j < PV; j++)
                if (cand[j]) { wlo = min(wlo, i0[j]); whi = max(whi, i0[j] + cand[j]); }
            wlo = __reduce_min_sync(0xFFFFFFFFu, wlo);
            whi = __reduce_max_sync(0xFFFFFFFFu, whi);
            // (only a compact window: the rows of an unsorted probe table point all over the build side, and walking from the
            //  smallest to the largest candidate would sweep the whole array per chunk -- 108 ms per 10M shuffled probes)
            if (whi > wlo && whi - wlo > 4u * (uint32_t)PF_CAP) whi = wlo;
            for (uint32_t e = (wlo & ~31u) + lane * 32u; e < whi; e += 32u * 32u) {  // one 128-byte line per lane and step
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ix.s_end + e));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ix.s_pos + e));
            }
        }
#endif
        uint32_t incl = lane_tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        uint32_t pj[PV];
        pj[0] = incl - lane_tot;
#pragma unroll
        for (int j = 1; j < PV; j++) pj[j] = pj[j - 1] + cand[j - 1];
        __syncwarp();  // the previous chunk's readers are done with the shared arrays
#pragma unroll
        for (int j = 0; j < PV; j++) sm.row[lane * (PV + 1) + j] = make_uint4(pj[j], i0[j], (uint32_t)qs[j], 0u);  // = flat_row_slot(lane * PV + j)
        if (lane == 31) sm.row[flat_row_slot(TP_WTILE)] = make_uint4(incl, 0u, 0u, 0u);
        __syncwarp();
        uint32_t* gb = build_idx + base;
        uint32_t* gp = probe_idx + base;
        const uint32_t row0 = (uint32_t)(chunk * TP_WTILE);
        uint32_t run = 0;  // pairs of this chunk written so far (warp-uniform)
        uint32_t ra = 0;   // rows [0, ra) are done (warp-uniform)
        while (ra < (uint32_t)TP_WTILE) {
            const uint32_t Pa = sm.row[flat_row_slot(ra)].x;
            // rows [ra, rb) whose candidates fit the slot map together (their prefix is monotone)
            uint32_t fit = 0;
#pragma unroll
            for (int j = 0; j < PV; j++) fit += (lane * PV + j >= ra && pj[j] + cand[j] - Pa <= (uint32_t)PF_CAP) ? 1u : 0u;
            fit = __reduce_add_sync(0xFFFFFFFFu, fit);
            if (fit == 0) {
                // one row alone exceeds the map: the warp sweeps its range 32 candidates at a time
                const uint32_t rI0 = sm.row[flat_row_slot(ra)].y, rN = sm.row[flat_row_slot(ra + 1)].x - Pa;
                const int32_t rQ = (int32_t)sm.row[flat_row_slot(ra)].z;
                for (uint32_t sb = 0; sb < rN; sb += 32) {
                    const uint32_t s = sb + lane;
                    const bool m = s < rN && __ldg(ix.s_end + rI0 + s) >= rQ;
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, m);
                    if (m) {
                        const uint32_t o = run + __popc(bal & lt_mask);
                        __stcs(gb + o, __ldg(ix.s_pos + rI0 + s));
                        __stcs(gp + o, row0 + ra);
                    }
                    run += __popc(bal);
                }
                ra++;
                continue;
            }
            const uint32_t rb = ra + fit;
            const uint32_t n_slots = sm.row[flat_row_slot(rb)].x - Pa;
            if (n_slots) {
                // ---- slot -> row map: zero, one byte per non-empty row at its first slot, packed prefix-max ---------
                const uint32_t n_words = (n_slots + 3) >> 2;
                uint32_t* w32 = reinterpret_cast<uint32_t*>(sm.rowof);
                for (uint32_t x = lane; x < n_words; x += 32) w32[x] = 0u;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < PV; j++) {
                    const uint32_t r = lane * PV + j;
                    if (r >= ra && r < rb && cand[j]) sm.rowof[pj[j] - Pa] = (uint8_t)r;
                }
                __syncwarp();
                const uint32_t per = ((n_words + 31) >> 5) | 1u;  // odd word stride per lane: conflict-free
                uint32_t carry = 0;
                for (uint32_t x = lane * per, e = min(lane * per + per, n_words); x < e; x++) {
                    uint32_t v = w32[x];
                    v = __vmaxu4(v, v << 8);
                    v = __vmaxu4(v, v << 16);
                    v = __vmaxu4(v, carry * 0x01010101u);
                    w32[x] = v;
                    carry = v >> 24;
                }
                uint32_t before = carry;  // exclusive prefix-max of the lanes' last bytes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, before, o);
                    if (lane >= (uint32_t)o) before = max(before, y);
                }
                before = __shfl_up_sync(0xFFFFFFFFu, before, 1);
                if (lane == 0) before = 0;
                if (before) {
                    const uint32_t bb = before * 0x01010101u;
                    for (uint32_t x = lane * per, e = min(lane * per + per, n_words); x < e; x++) w32[x] = __vmaxu4(w32[x], bb);
                }
                __syncwarp();
                // ---- test 32 candidate slots per step, compact, write one contiguous run ---------------------------
                // (PF_U steps are software-pipelined: all their loads are issued before the first ballot)
                for (uint32_t sb = 0; sb < n_slots; sb += 32 * PF_U) {
                    uint32_t rr[PF_U], pos[PF_U];
                    bool mt[PF_U];
#pragma unroll
                    for (int u = 0; u < PF_U; u++) {
                        const uint32_t s = sb + u * 32 + lane;
                        const bool v = s < n_slots;
                        rr[u] = v ? sm.rowof[s] : ra;
                        const uint4 rw = sm.row[flat_row_slot(rr[u])];
                        const uint32_t i = rw.y + (s + Pa - rw.x);
                        mt[u] = v && __ldg(ix.s_end + i) >= (int32_t)rw.z;
                        pos[u] = v ? __ldg(ix.s_pos + i) : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < PF_U; u++) {
                        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, mt[u]);
                        if (mt[u]) {
                            const uint32_t o = run + __popc(bal & lt_mask);
                            __stcs(gb + o, pos[u]);
                            __stcs(gp + o, row0 + rr[u]);
                        }
                        run += __popc(bal);
                    }
                }
            }
            ra = rb;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Output materialisation on the device: arrow::compute::take of a fixed-width column by the emitted index column
// (interval_join.rs:1655-1667, nearest.rs:456-482).  One element per thread; the index column of a probe-major pair list
// is non-decreasing on the probe side and locally clustered on the build side, so the gathers hit L1/L2.
// NULL indices (validity bit clear) produce zeroed elements; the output validity is the index validity.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void take_kernel(const T* __restrict__ src, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ valid_bits,
                            uint64_t n, T* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool ok = !valid_bits || ((__ldg(valid_bits + (i >> 5)) >> (i & 31)) & 1u);
        T v{};
        if (ok) v = __ldg(src + __ldcs(idx + i));
        __stcs(out + i, v);
    }
}
int launch_take(const void* src, uint32_t elem_bytes, const uint32_t* idx, const uint32_t* valid_bits, uint64_t n, void* out,
                cudaStream_t st) {
    if (n == 0) return 0;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)kSMs * 32));
    switch (elem_bytes) {
        case 1: take_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)src, idx, valid_bits, n, (uint8_t*)out); break;
        case 2: take_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)src, idx, valid_bits, n, (uint16_t*)out); break;
        case 4: take_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)src, idx, valid_bits, n, (uint32_t*)out); break;
        case 8: take_kernel<unsigned long long><<<grid, 256, 0, st>>>((const unsigned long long*)src, idx, valid_bits, n, (unsigned long long*)out); break;
        case 16: take_kernel<uint4><<<grid, 256, 0, st>>>((const uint4*)src, idx, valid_bits, n, (uint4*)out); break;
        default: set_error("brg_take: element width must be 1, 2, 4, 8 or 16 bytes"); return 1;
    }
    BRG_LAUNCH_CHECK();
    return 0;
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
    // A build side with at least as many rows as the probe table puts more than TP_WMAX elements under most chunks (unless the
    // probe table covers only a slice of the build side's coordinate range, which the row counts cannot tell): the tiled
    // kernel would look at each warp's first chunk, find its window too wide and hand everything over (19 us per 10M rows
    // with the to-do list).  Such calls go to the one-row-per-thread kernel directly; both paths are exact, this only picks
    // the likely faster one.  (Not applied to count / coverage / nearest: their host-buffer entry points probe in 8M-row
    // slices of a sorted table, where a 10M-row build side is sparse.)
    const bool dense = (uint64_t)ix.n >= n;
    p.tiled = pairs_tiled_enabled() && !dense && p.full_rows > 0 && aligned16p(key) && aligned16p(s) && aligned16p(e) &&
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
                               uint64_t* chunk_base, uint64_t* row_rng, cudaStream_t st) {
    const ChunkPlan p = plan_chunks(ix, key, s, e, n);
    if (p.n_chunks == 0) { BRG_CUDA(cudaMemsetAsync(chunk_base, 0, sizeof(uint64_t), st)); return 0; }
    // per-call temporaries from the stream-ordered pool; one exit frees them whatever happened in between
    uint64_t* tot = nullptr;
    uint64_t* scratch = nullptr;
    TodoList todo{nullptr, nullptr};
    auto body = [&]() -> int {
        int rc = 0;
        BRG_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tot), p.n_chunks * sizeof(uint64_t), st));
        BRG_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), (scan_scratch_words(p.n_chunks) + 16) * 4, st));
        if (p.tiled) {
            if ((rc = todo_list_alloc(p.n_chunks, &todo, st))) return rc;
            BRG_CUDA(cudaFuncSetAttribute((const void*)pairs_ranges_kernel_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProbeRing)));
            const uint64_t tiles = p.full_rows / TP_TILE + 1;
            const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)kSMs * BRG_PR_CTAS));
            pairs_ranges_kernel_tiled<<<grid, TP_THREADS, sizeof(ProbeRing), st>>>(ix, key, s, e, p.full_rows, row_rng, tot, todo);
            BRG_LAUNCH_CHECK();
            pairs_ranges_todo_kernel<<<chunk_todo_grid(p.n_chunks), TP_WTILE, 0, st>>>(ix, key, s, e, n, row_rng, tot, todo, 0u, p.tail_chunk);
            BRG_LAUNCH_CHECK();
        } else {
            pairs_ranges_todo_kernel<<<chunk_todo_grid(p.n_chunks), TP_WTILE, 0, st>>>(ix, key, s, e, n, row_rng, tot, TodoList{nullptr, nullptr},
                                                                                   (uint32_t)p.n_chunks, 0xFFFFFFFFu);
            BRG_LAUNCH_CHECK();
        }
        return exclusive_scan_u64(tot, chunk_base, p.n_chunks, scratch, st);
    };
    const int rc = body();
    if (todo.count) cudaFreeAsync(todo.count, st);
    if (tot) cudaFreeAsync(tot, st);
    if (scratch) cudaFreeAsync(scratch, st);
    return rc;
}

int launch_pairs_fill(const IndexView& ix, const int32_t* s, uint64_t n, const uint64_t* chunk_base, const uint64_t* row_rng,
                      uint32_t* build_idx, uint32_t* probe_idx, cudaStream_t st) {
    const uint64_t n_chunks = pairs_chunk_count(n);
    if (n_chunks == 0) return 0;
    BRG_CUDA(cudaFuncSetAttribute((const void*)pairs_emit_kernel_flat, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(FlatWarp) * PF_WARPS)));
    const uint64_t ctas = (n_chunks + PF_WARPS - 1) / PF_WARPS;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(ctas, (uint64_t)kSMs * BRG_PF_CTAS));
    const int vec_ok = aligned16p(s) && aligned16p(row_rng);
    pairs_emit_kernel_flat<<<grid, PF_THREADS, sizeof(FlatWarp) * PF_WARPS, st>>>(ix, s, n, n_chunks, row_rng, chunk_base, build_idx,
                                                                                 probe_idx, vec_ok);
    BRG_LAUNCH_CHECK();
    return 0;
}

}  // namespace brg
