// brg_internal.cuh -- shared declarations of the sm_100a interval engine (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "brg_bins.cuh"

namespace brg {

constexpr int kSMs = 148;  // B200: 148 SMs (2 dies x 74); grids are sized in multiples of this

// ------------------------------------------------------------------------------------------------
// error plumbing: no exceptions cross the ABI; the message is thread-local
// ------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launches;

#define BRG_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            brg::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
            return 2; /* BRG_ERR_CUDA */                                                            \
        }                                                                                           \
    } while (0)

#define BRG_LAUNCH_CHECK()                                                                          \
    do {                                                                                            \
        brg::g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess) {                                                                    \
            brg::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));                \
            return 2;                                                                               \
        }                                                                                           \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Rank directory: one level of direct addressing in front of a per-key sorted i32 array A.
// For key c the value range [lo, hi] is cut into nb bins of width 2^shift; D[base + j] is the global
// index of the first element of the key's segment whose bin is >= j, D[base + nb] = seg_end.
// rank_le(q) therefore needs one KeyDir sector, two adjacent D words and a search over one bin.
// ------------------------------------------------------------------------------------------------
struct __align__(16) KeyDir {
    int32_t lo;         // smallest value of the segment
    uint32_t nb;        // number of bins (0 when the segment is empty)
    uint32_t base;      // first D entry of this key
    uint32_t seg_begin; // segment [seg_begin, seg_end) in A
    uint32_t seg_end;
    uint32_t pad[3];    // 32 B = one L2 sector
};

struct DirView {
    const KeyDir* kd;
    const uint32_t* D;
    const int32_t* A;
    uint32_t shift;
};

constexpr uint32_t BRG_IDX_INVERTED = 1u << 8;  // some build row has start > end (count identity does not hold for pairs)

// any row with start > end?  (sets *flag to 1)
int flag_inverted_rows(const int32_t* start, const int32_t* end, uint64_t n, uint32_t* d_flag, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// Compact rank index: the search structure of the one-row-per-thread kernels (unsorted probe tables).
// A random rank query against the plain arrays costs ~8 divergent load instructions (KeyDir, two directory words, a
// binary search over a bin of A) and touches sectors that do not stay in L2 next to the probe stream
// (profiles/r1_count_todo_v1_ncu.txt: 9.9 GB of DRAM reads for 1.5 GB of probe columns).  Fully divergent loads are
// paid per lane in the L1 pipeline, so the structure below is built to need THREE of them per query:
//   blocks : every key segment cut into runs of 7 consecutive elements, 16 bytes each:
//            { int32 first element F ; u16 delta[6] of the next 6 elements from F }
//            delta[0] == 0xFFFF marks a block whose deltas do not fit (the query then searches A itself);
//            missing tail elements are 0xFFFE, above every clamped query; real deltas are <= 0xFFFD.
//   P      : direct-address table over the VALUES: P[bin] = (first block starting in the bin) << 2 | min(#block starts
//            in the bin, 3), about one bin per block.
//   keys   : the per-key constants, copied to shared memory by the kernels when they fit.
// Query = P[bin] -> the block before the bin, or the (1-2) blocks starting in it -> 16-byte block -> SIMD compare.
// KeyDir of `d`: seg_begin/seg_end are block indices; pad[0]/pad[1] hold the ELEMENT range of the key in A.
// ------------------------------------------------------------------------------------------------
constexpr int CI_BLOCK = 7;
struct CompactView {
    DirView d;            // d.kd = per-key constants, d.D = packed table P, d.shift = bin width; d.A unused
    const uint4* blocks;  // one uint4 per block; nullptr = structure absent
    const int32_t* A;     // the plain sorted array (wide blocks)
};
struct CompactKey {       // per-key constants of a CompactView (32 bytes, same layout as KeyDir)
    int32_t lo;
    uint32_t nb, base, blk_begin, blk_end, eb, ee, pad;
};

// Bins over a PAIR of sorted arrays that share their key segments (see "Count bins" below)
struct BinsView {
    const uint4* rec;   // 2 x uint4 per bin; nullptr = absent
    const KeyDir* kd;   // per key: lo, nb, base (first record), seg_begin / seg_end (shared by the two arrays)
    uint32_t shift;     // bin width 2^shift (format 0: <= 65536, in-bin offsets are u16; format 1: <= 1024, 10-bit offsets)
    uint32_t fmt;       // record format (brg_bins.cuh): 0 = 12 x u16 slots, 1 = 18 x 10-bit slots
};

// Device view of an index (pointers into the single HBM slab)
struct IndexView {
    uint32_t n_keys;
    uint32_t n;        // build rows with a valid key
    uint32_t n_merged; // merged (coverage) intervals
    uint32_t flags;    // BRG_WANT_* actually built | BRG_IDX_* properties
    // S-order: rows sorted by (key, start, end, pos)
    const int32_t* s_start;
    const int32_t* s_end;
    const uint32_t* s_pos;
    const int32_t* pme; // per-key inclusive running max of s_end  (nearest_index.rs:58-63)
    // E-order: rows sorted by (key, end, start, pos)
    const int32_t* e_end;
    // merged intervals per key (interval_tree.rs:52-73) + exclusive prefix of widths max(1,last-first)
    const int32_t* m_first;
    const int32_t* m_last;
    const int32_t* m_pre;
    // 16-byte records for the epilogues (one 128-bit load per candidate instead of three or four 32-bit ones):
    //   s_rec[i] = {start, end, pos, prefix_max_end}  in S-order      (nearest, nearest_k)
    //   e_rec[i] = {start, end, pos, 0}               in E-order      (nearest: left neighbour)
    //   m_rec[j] = {first, last, pre[j], pre[j+1]}    merged          (coverage)
    const int4* s_rec;
    const int4* e_rec;
    const int4* m_rec;
    DirView d_sstart, d_eend, d_pme, d_mfirst, d_mlast;
    CompactView c_sstart, c_eend;  // per-row path of count_overlaps (superseded by the count bins when those are present)
    // Count bins: ONE 32-byte record per coordinate bin answers both rank queries of a count_overlaps row (see CountBin)
    BinsView cb;   // over (s_start, e_end): count_overlaps
    BinsView cvb;  // over (m_first, m_last): coverage -- the same record answers "first merged interval with last >= qs" and
                   // "first merged interval with first > qe"
    BinsView nb;   // over (s_start, pme): nearest and the pair range pass on unsorted probe tables -- one record answers
                   // #(start <= qe) and "first row whose prefix_max_end >= qs" (absent when a build row has start > end)
};

// ------------------------------------------------------------------------------------------------
// Count bins: the search structure of count_overlaps for UNSORTED probe tables.  A shuffled probe row costs one fully
// divergent load instruction per structure it touches, and the L1 pipeline pays those per lane: the per-row kernels run at
// about 32 cycles per divergent load instruction and warp whatever its width (measured: 5-6 of them per row through the
// compact rank index = 3.2 ms per 125M rows, whether the index has 1M or 10M rows; ncu: long_scoreboard 53 warps per issue,
// lg_throttle 8 -- the load/store queue is the bound, not DRAM, not L2 capacity, not issue slots).  So the bins fold everything
// a row needs into ONE 32-byte record = one sector = two 128-bit loads:
//   bin = (coordinate - key.lo) >> shift, about 2 starts + 2 ends per bin
//   record = { cs : global index of the first start >= the bin's first coordinate                          (u32)
//              depth : cs - (index of the first end >= the bin's first coordinate) = intervals open there   (u16)
//              ns, ne : starts / ends inside the bin                                                        (u8, u8)
//              12 x u16 : in-bin offsets of the ns starts, then of the ne ends, both ascending }
//   ne == 0xFF marks a record that does not fit (more than 12 elements, depth >= 65535): the row searches the plain arrays.
//   It has to be rare -- one overflowing lane sends its whole warp through the slow path -- hence the narrow bins.
//   #(starts <= qe) = cs(bin(qe)) + #(start offsets <= qe's offset);   #(ends < qs) likewise from bin(qs - 1).
// A 150 bp read has both coordinates in one bin most of the time: 2 divergent loads per row instead of 5-6.
// ------------------------------------------------------------------------------------------------
// (the record layout and its two formats: brg_bins.cuh)

// ------------------------------------------------------------------------------------------------
// launchers (each returns a brg_status; all work is enqueued on `st`)
// ------------------------------------------------------------------------------------------------
// K1 radix sort of rows (SoA of up to 4 u32 words per row)
struct SortPass {
    int src;          // which word array supplies the digit
    int shift;        // bit offset of the 8-bit digit
    int flip_sign;    // 1: value is i32, order as signed (flip the sign bit)
    uint32_t clamp;   // values above clamp are treated as clamp (invalid keys sort last); 0xFFFFFFFF = off
};
// sorts n rows; in/out are ping-pong sets of `n_arr` u32 arrays; returns which set holds the result in *res
int radix_sort_rows(uint32_t* const* a, uint32_t* const* b, int n_arr, uint64_t n, const SortPass* passes,
                    int n_passes, uint32_t* d_hist_scratch, uint64_t hist_scratch_words, uint32_t* d_scan_scratch,
                    cudaStream_t st, int* res);
uint64_t radix_hist_words(uint64_t n);

// K2 scans
uint64_t scan_scratch_words(uint64_t n);
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* scratch, cudaStream_t st);
int exclusive_scan_u32_to_u64(const uint32_t* in, uint64_t* out /*n+1*/, uint64_t n, uint64_t* scratch,
                              cudaStream_t st);
int exclusive_scan_u64(const uint64_t* in, uint64_t* out /*n+1*/, uint64_t n, uint64_t* scratch, cudaStream_t st);
int segmented_running_max(const uint32_t* key, const int32_t* val, int32_t* out, uint64_t n, int32_t* scratch,
                          cudaStream_t st);
// merge_intervals on the S-order arrays: writes m_first/m_last/(m_key) and *d_n_merged
int merge_sorted_intervals(const uint32_t* s_key, const int32_t* s_start, const int32_t* pme, uint64_t n,
                           uint32_t* d_flags /*n*/, uint32_t* d_midx /*n*/, uint32_t* scan_scratch, int32_t* m_first,
                           int32_t* m_last, uint32_t* m_key, uint32_t* d_n_merged, cudaStream_t st);
// merge UDTF: runs of rows sorted by (key, start, end); writes run heads and *d_n_merged, then the row counts
int merge_runs(const uint32_t* s_key, const int32_t* s_start, const int32_t* pme, uint64_t n, long long min_dist, int strict,
               uint32_t* d_flags, uint32_t* d_midx, uint32_t* scan_scratch, uint32_t* m_key, int32_t* m_first, int32_t* m_last,
               uint32_t* m_row0, uint32_t* d_n_merged, cudaStream_t st);
int merge_counts(const uint32_t* m_row0, uint64_t m, uint64_t n, uint32_t* m_count, cudaStream_t st);
int cluster_rows(const uint32_t* flags, const uint32_t* midx, const int32_t* m_first, const int32_t* m_last, uint64_t n,
                 uint32_t* cluster, int32_t* c_start, int32_t* c_end, cudaStream_t st);
// complement UDTF: per-view row counts + plan (count pass), then one output row per thread (emit pass)
int complement_plan(const uint32_t* m_key, const int32_t* m_first, const int32_t* m_last, uint64_t m, uint32_t n_keys,
                    const uint32_t* v_key, const long long* v_start, const long long* v_end, uint64_t n_views,
                    uint32_t* m_off /*n_keys+1*/, uint32_t* g_tmp /*m*/, uint32_t* G /*m+1*/, uint32_t* scan_scratch,
                    uint4* plan /*n_views*/, uint32_t* counts /*n_views*/, cudaStream_t st);
int complement_emit(const long long* v_start, const long long* v_end, uint64_t n_views, const uint64_t* offsets,
                    const int32_t* m_first, const int32_t* m_last, const uint32_t* G, const uint4* plan, uint64_t total,
                    uint32_t* out_view, long long* out_start, long long* out_end, cudaStream_t st);
int merged_width_prefix(const int32_t* m_first, const int32_t* m_last, uint64_t n_merged, uint32_t* widths_tmp,
                        int32_t* m_pre /*n_merged+1*/, uint32_t* scan_scratch, cudaStream_t st);
// segment offsets off[0..n_keys] from a sorted key array (keys >= n_keys are the invalid tail)
int segment_offsets(const uint32_t* sorted_key, uint64_t n, uint32_t n_keys, uint32_t* off, cudaStream_t st);
// 16-byte records: out[i] = {a[i], b[i], c[i], d ? d[i] : 0}
int pack_records(const int32_t* a, const int32_t* b, const int32_t* c, const int32_t* d, uint64_t n, int4* out,
                 cudaStream_t st);
// compact rank index: S[j] / blocks[j] of every 16-element block; boff = per-key block offsets (n_keys + 1)
int ci_build(const int32_t* A, const uint32_t* off, const uint32_t* boff, uint32_t n_keys, uint32_t n_blocks, int32_t* S,
             uint4* blocks, cudaStream_t st);
int ci_patch_keydir(KeyDir* kd, const uint32_t* off, uint32_t n_keys, cudaStream_t st);
// P[i] = D[i] << 2 | min(D[i + 1] - D[i], 3)
int ci_pack_table(const uint32_t* D, uint32_t total_entries, uint32_t* P, cudaStream_t st);
// count bins: per-key value range over BOTH arrays (kd.nb = provisional range), then one thread per record
int cb_key_ranges(const int32_t* s_start, const int32_t* e_end, const uint32_t* off, uint32_t n_keys, KeyDir* kd,
                  unsigned long long* d_bins_by_shift, cudaStream_t st);
int cb_fill(const DirView& d_sstart, const DirView& d_eend, const KeyDir* kd, uint32_t n_keys, uint32_t shift, uint32_t fmt, uint32_t* cs_tmp,
            uint32_t* ce_tmp /* total_entries + 1 words each */, uint4* rec, uint32_t total_entries, cudaStream_t st);
// rank directory
int dir_key_ranges(const int32_t* A, const uint32_t* off, uint32_t n_keys, KeyDir* kd, unsigned long long* d_bins_by_shift,
                   cudaStream_t st);
int dir_assign_bases(KeyDir* kd, uint32_t n_keys, uint32_t shift, uint32_t* d_total_entries, cudaStream_t st);
int dir_fill(const int32_t* A, const KeyDir* kd, uint32_t n_keys, uint32_t shift, uint32_t* D, uint32_t total_entries,
             cudaStream_t st);

// K3 probes
int launch_count(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n, int strict,
                 int64_t* out, cudaStream_t st);
int launch_coverage(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                    int strict, int64_t* out, cudaStream_t st);
int launch_nearest_k1(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                      int include_overlaps, int strict, uint32_t* left_idx, uint32_t* valid_bits, int64_t* dist,
                      cudaStream_t st);
// subtract: offsets == nullptr -> count pass (counts), else emit pass
int launch_subtract(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n, int strict,
                    uint32_t* counts, const uint64_t* offsets, uint32_t* out_row, int32_t* out_start, int32_t* out_end,
                    cudaStream_t st);
int launch_nearest_k_count(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                           uint32_t k, int include_overlaps, int strict, uint32_t* counts, cudaStream_t st);
int launch_nearest_k_emit(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                          uint32_t k, int include_overlaps, int strict, const uint64_t* offsets, uint32_t* build_idx,
                          uint32_t* valid_bytes, uint32_t* probe_idx, int64_t* dist, cudaStream_t st);
// K4 pairs
int launch_pairs_count(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                       uint32_t* counts, cudaStream_t st);
int launch_pairs_emit(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                      const uint64_t* offsets, uint32_t* build_idx, uint32_t* probe_idx, cudaStream_t st);
// chunked pair pipeline: chunk_base[c] = first output slot of 256-row chunk c, chunk_base[n_chunks] = total
uint64_t pairs_chunk_count(uint64_t n);
int launch_pairs_chunk_offsets(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n,
                               uint64_t* chunk_base /*n_chunks + 1*/, uint64_t* row_rng /*n*/, cudaStream_t st);
int launch_pairs_fill(const IndexView& ix, const int32_t* s, uint64_t n, const uint64_t* chunk_base, const uint64_t* row_rng,
                      uint32_t* build_idx, uint32_t* probe_idx, cudaStream_t st);
// the pair join in ONE launch (brg_join.cu): ranges -> count -> output slots -> emit.  Pairs whose output slot is >= capacity are
// dropped (the total stays exact); d_state = pairs_join_state_bytes(n) bytes of scratch; chunk_slots_out may be NULL
uint64_t pairs_join_state_bytes(uint64_t n);
uint32_t pairs_join_chunk_rows();
int launch_pairs_join(const IndexView& ix, const uint32_t* key, const int32_t* s, const int32_t* e, uint64_t n, uint64_t capacity,
                      int ordered, uint32_t* build_idx, uint32_t* probe_idx, uint64_t* d_total, uint64_t* chunk_slots_out, void* d_state,
                      cudaStream_t st);
int launch_take(const void* src, uint32_t elem_bytes, const uint32_t* idx, const uint32_t* valid_bits, uint64_t n, void* out,
                cudaStream_t st);
int launch_semi_anti_flags(const uint32_t* counts, uint64_t n, int anti, uint32_t* flags, cudaStream_t st);
int launch_compact_rows(const uint32_t* flags, const uint64_t* offsets, uint64_t n, uint32_t* probe_idx,
                        cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ KeyDir load_keydir(const KeyDir* kd, uint32_t c) {
    const int4* p = reinterpret_cast<const int4*>(kd + c);
    int4 a = __ldg(p);
    int4 b = __ldg(p + 1);
    KeyDir k;
    k.lo = a.x; k.nb = (uint32_t)a.y; k.base = (uint32_t)a.z; k.seg_begin = (uint32_t)a.w; k.seg_end = (uint32_t)b.x;
    return k;
}

// global index of the first element of key c's segment that is > q  (== seg_begin + #(A <= q))
__device__ __forceinline__ uint32_t rank_le_kd(const DirView& d, const KeyDir& k, int32_t q) {
    if (k.nb == 0 || q < k.lo) return k.seg_begin;
    uint32_t b = ((uint32_t)q - (uint32_t)k.lo) >> d.shift;
    if (b >= k.nb) return k.seg_end;
    uint32_t lo = __ldg(d.D + k.base + b), hi = __ldg(d.D + k.base + b + 1);
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(d.A + mid) <= q) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// The directory step of rank_le_kd alone: [lo, hi) = the part of the key's segment that still has to be searched for q
// (lo == hi when the directory already decides the rank).  Several independent rank queries of one row can take this
// step together and then run their bin searches in lock step (rank_search3), so their memory latencies overlap.
__device__ __forceinline__ void rank_le_bin(const DirView& d, const KeyDir& k, int32_t q, uint32_t* lo, uint32_t* hi) {
    if (k.nb == 0 || q < k.lo) { *lo = *hi = k.seg_begin; return; }
    const uint32_t b = ((uint32_t)q - (uint32_t)k.lo) >> d.shift;
    if (b >= k.nb) { *lo = *hi = k.seg_end; return; }
    *lo = __ldg(d.D + k.base + b);
    *hi = __ldg(d.D + k.base + b + 1);
}
__device__ __forceinline__ void rank_lt_bin(const DirView& d, const KeyDir& k, int32_t q, uint32_t* lo, uint32_t* hi) {
    if (q == INT32_MIN) { *lo = *hi = k.seg_begin; return; }
    rank_le_bin(d, k, q - 1, lo, hi);
}
// three bin searches in lock step: on return lo_i = rank (first element of A_i in [lo_i, hi_i) that is > q_i)
__device__ __forceinline__ void rank_search3(const int32_t* A0, int32_t q0, uint32_t& lo0, uint32_t hi0, const int32_t* A1,
                                             int32_t q1, uint32_t& lo1, uint32_t hi1, const int32_t* A2, int32_t q2,
                                             uint32_t& lo2, uint32_t hi2) {
    while ((lo0 < hi0) | (lo1 < hi1) | (lo2 < hi2)) {
        const uint32_t m0 = (lo0 + hi0) >> 1, m1 = (lo1 + hi1) >> 1, m2 = (lo2 + hi2) >> 1;
        const bool a0 = lo0 < hi0, a1 = lo1 < hi1, a2 = lo2 < hi2;
        const int32_t v0 = a0 ? __ldg(A0 + m0) : 0, v1 = a1 ? __ldg(A1 + m1) : 0, v2 = a2 ? __ldg(A2 + m2) : 0;
        if (a0) { if (v0 <= q0) lo0 = m0 + 1; else hi0 = m0; }
        if (a1) { if (v1 <= q1) lo1 = m1 + 1; else hi1 = m1; }
        if (a2) { if (v2 <= q2) lo2 = m2 + 1; else hi2 = m2; }
    }
}
// global index of the first element that is >= q (== seg_begin + #(A < q))
__device__ __forceinline__ uint32_t rank_lt_kd(const DirView& d, const KeyDir& k, int32_t q) {
    if (q == INT32_MIN) return k.seg_begin;
    return rank_le_kd(d, k, q - 1);
}

// seg_begin(A) + #(A <= q) within a key, through the compact index; k = the key's constants (shared or global memory)
// (the delta blocks are loaded with ld.global.cg: every lane reads a different one, an L1 line allocation per lane buys nothing
//  -- shuffled 125M-probe count 3.35 -> 3.26 ms; the packed table keeps ld.global.nc, its hot entries do hit in L1)
__device__ __forceinline__ uint32_t rank_le_compact(const CompactView& c, const CompactKey& k, int32_t q) {
    if (k.nb == 0 || q < k.lo) return k.eb;
    // J = last block whose first element is <= q
    const uint32_t bin = ((uint32_t)q - (uint32_t)k.lo) >> c.d.shift;
    uint32_t J;
    uint4 w;
    if (bin >= k.nb) {
        J = k.blk_end - 1;
        w = __ldcg(c.blocks + J);
    } else {
        const uint32_t e = __ldg(c.d.D + k.base + bin);
        const uint32_t j0 = e >> 2;
        uint32_t cnt = e & 3u;
        J = j0 - 1;  // (j0 > blk_begin whenever no block starting in this bin is <= q, see the builder)
        bool found = false;
        if (cnt) {
            if (cnt == 3u) cnt = (__ldg(c.d.D + k.base + bin + 1) >> 2) - j0;  // saturated: the real count
            for (uint32_t j = j0; j < j0 + cnt; j++) {  // block starts inside the bin: one, seldom more
                const uint4 t = __ldcg(c.blocks + j);
                if ((int32_t)t.x > q) break;
                J = j; w = t; found = true;
            }
        }
        if (!found) w = __ldcg(c.blocks + J);
    }
    const uint32_t e0 = k.eb + (J - k.blk_begin) * CI_BLOCK;
    if ((w.y & 0xFFFFu) == 0xFFFFu) {  // wide block: search the elements themselves
        uint32_t lo = e0, hi = min(e0 + CI_BLOCK, k.ee);
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(c.A + mid) <= q) lo = mid + 1; else hi = mid;
        }
        return lo;
    }
    const uint32_t t = min((uint32_t)q - w.x, 0xFFFDu);
    const uint32_t tt = t | (t << 16);
    const uint32_t bits = __popc(__vcmpleu2(w.y, tt)) + __popc(__vcmpleu2(w.z, tt)) + __popc(__vcmpleu2(w.w, tt));
    return e0 + 1 + (bits >> 4);  // the block's first element + 16 bits per matching halfword
}
__device__ __forceinline__ uint32_t rank_lt_compact(const CompactView& c, const CompactKey& k, int32_t q) {
    if (q == INT32_MIN) return k.eb;
    return rank_le_compact(c, k, q - 1);
}
__device__ __forceinline__ CompactKey load_compact_key(const CompactView& c, uint32_t key) {
    const int4* p = reinterpret_cast<const int4*>(c.d.kd + key);
    const int4 a = __ldg(p), b = __ldg(p + 1);
    CompactKey k;
    k.lo = a.x; k.nb = (uint32_t)a.y; k.base = (uint32_t)a.z; k.blk_begin = (uint32_t)a.w;
    k.blk_end = (uint32_t)b.x; k.eb = (uint32_t)b.y; k.ee = (uint32_t)b.z; k.pad = 0;
    return k;
}

#ifndef BRG_LD256
#define BRG_LD256 1
#endif
// One count bin record (32 bytes = one sector), loaded once and used for both rank queries of a row.
struct CountBinRec { uint32_t w[8]; };
__device__ __forceinline__ CountBinRec load_count_bin(const BinsView& bv, uint32_t rec) {
    const uint4* r = bv.rec + 2ull * rec;
    CountBinRec x;
#if BRG_LD256
    // ONE 256-bit load (sm_100: LDG.E.256) per 32-byte record: a shuffled probe row pays per divergent load INSTRUCTION
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(x.w[0]), "=r"(x.w[1]), "=r"(x.w[2]), "=r"(x.w[3]), "=r"(x.w[4]), "=r"(x.w[5]), "=r"(x.w[6]), "=r"(x.w[7])
                 : "l"(r));
#else
    const uint4 a = __ldg(r), b = __ldg(r + 1);
    x.w[0] = a.x; x.w[1] = a.y; x.w[2] = a.z; x.w[3] = a.w; x.w[4] = b.x; x.w[5] = b.y; x.w[6] = b.z; x.w[7] = b.w;
#endif
    return x;
}
__device__ __forceinline__ bool count_bin_overflowed(const CountBinRec& x) { return cb_overflowed(x.w); }
// #(elements of the bin <= off) among the starts (ENDS = false) or the ends (true) of a record that did not overflow
template <bool ENDS>
__device__ __forceinline__ uint32_t count_bin_slots(const CountBinRec& x, uint32_t fmt, uint32_t off) {
    return cb_count_le<ENDS>(x.w, fmt, off);
}
// global index of the first start / end >= the bin's first coordinate
__device__ __forceinline__ uint32_t count_bin_cs(const CountBinRec& x) { return cb_first_a(x.w); }
__device__ __forceinline__ uint32_t count_bin_ce(const CountBinRec& x) { return cb_first_b(x.w); }

// Both rank queries of a row through the bins of one array pair:
//   *ra = seg_begin + #(first array <= qa),  *rb = seg_begin + #(second array < qb)
// A rank whose record overflowed is left for the caller (returned bit 0: ra missing, bit 1: rb missing).
__device__ __forceinline__ uint32_t bins_pair_ranks(const BinsView& bv, const KeyDir& k, int32_t qa, int32_t qb, uint32_t* ra,
                                                    uint32_t* rb) {
    // bin of qa, bin of qb - 1; outside [lo, lo + nb * width) the segment bounds answer
    const bool a_in = qa >= k.lo, b_in = qb > k.lo;
    const uint32_t da = (uint32_t)qa - (uint32_t)k.lo, db = (uint32_t)qb - 1u - (uint32_t)k.lo;
    const uint32_t ba = da >> bv.shift, bb = db >> bv.shift;
    const bool need_a = a_in && ba < k.nb, need_b = b_in && bb < k.nb;
    *ra = a_in ? k.seg_end : k.seg_begin;
    *rb = b_in ? k.seg_end : k.seg_begin;
    const uint32_t msk = (1u << bv.shift) - 1u;
    uint32_t missing = 0;
    CountBinRec r;
    if (need_a) {
        r = load_count_bin(bv, k.base + ba);
        if (count_bin_overflowed(r)) missing |= 1u;
        else *ra = count_bin_cs(r) + count_bin_slots<false>(r, bv.fmt, da & msk);
    }
    if (need_b) {
        if (!(need_a && bb == ba)) r = load_count_bin(bv, k.base + bb);  // (the usual case: the row lies inside one bin)
        if (count_bin_overflowed(r)) missing |= 2u;
        else *rb = count_bin_ce(r) + count_bin_slots<true>(r, bv.fmt, db & msk);
    }
    return missing;
}

__device__ __forceinline__ int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

// streaming (touch-once) loads/stores: keep the probe stream out of L1 and first in line for L2 eviction so the
// index stays resident in the 126 MB L2
__device__ __forceinline__ int32_t ld_stream(const int32_t* p) { return __ldcs(p); }
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(int64_t* p, int64_t v) { __stcs(reinterpret_cast<long long*>(p), (long long)v); }
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v) { __stcs(p, v); }
#endif

}  // namespace brg
