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
#include "brg_internal.cuh"

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

}  // namespace brg
