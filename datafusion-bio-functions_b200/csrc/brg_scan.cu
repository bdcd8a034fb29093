// brg_scan.cu -- K2: device-wide scans and the build-side post-processing that runs on them (sm_100a).
//
//  * exclusive prefix sums (radix-sort histogram table, merged-interval indices, width prefix, pair offsets)
//  * per-key segmented running max of `end` over the start-sorted rows  == prefix_max_end
//    (nearest_index.rs:58-63) and the accumulator of merge_intervals (interval_tree.rs:52-73)
//  * merge_intervals itself as flag -> scan -> compact, and the width prefix used by coverage
//  * segment offsets per key and the rank directories (see brg_internal.cuh)
//
// Scans are reduce-then-scan with warp-shuffle block scans: tile reduce -> one-CTA scan of the tile
// aggregates -> tile scan with carry-in.  All of it is HBM-bound streaming (2 reads + 1 write per element).
#include <climits>

#include "brg_internal.cuh"

namespace brg {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;  // 2048

// ---------------------------------------------------------------------------------------------------
// operators
// ---------------------------------------------------------------------------------------------------
struct SegMaxT {
    int32_t v;
    uint32_t f;  // 1 when a segment head lies inside the span this value summarises
};
template <typename T>
struct SumOp {
    __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
    __device__ __forceinline__ static T identity() { return T(0); }
};
struct SegMaxOp {
    __device__ __forceinline__ SegMaxT operator()(SegMaxT a, SegMaxT b) const {
        SegMaxT r;
        r.v = b.f ? b.v : max(a.v, b.v);
        r.f = a.f | b.f;
        return r;
    }
    __device__ __forceinline__ static SegMaxT identity() { return SegMaxT{INT32_MIN, 0u}; }
};

__device__ __forceinline__ uint32_t shfl_up_t(uint32_t v, int o) { return __shfl_up_sync(0xFFFFFFFFu, v, o); }
__device__ __forceinline__ uint64_t shfl_up_t(uint64_t v, int o) { return __shfl_up_sync(0xFFFFFFFFu, v, o); }
__device__ __forceinline__ SegMaxT shfl_up_t(SegMaxT v, int o) {
    SegMaxT r;
    r.v = __shfl_up_sync(0xFFFFFFFFu, v.v, o);
    r.f = __shfl_up_sync(0xFFFFFFFFu, v.f, o);
    return r;
}

// inclusive scan of one value per thread across the CTA; returns the inclusive value, *block_total = aggregate
template <typename T, typename Op, int THREADS>
__device__ __forceinline__ T block_inclusive_scan(T x, Op op, T* smem /*THREADS/32*/, T* block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T y = shfl_up_t(x, o);
        if (lane >= o) x = op(y, x);
    }
    if (lane == 31) smem[w] = x;
    __syncthreads();
    T carry = Op::identity();
    T tot = Op::identity();
#pragma unroll
    for (int ww = 0; ww < THREADS / 32; ww++) {
        T s = smem[ww];
        if (ww < w) carry = op(carry, s);
        tot = op(tot, s);
    }
    __syncthreads();
    *block_total = tot;
    return op(carry, x);
}

// ---------------------------------------------------------------------------------------------------
// loaders / storers
// ---------------------------------------------------------------------------------------------------
struct LoadU32 {
    const uint32_t* in;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return in[i]; }
};
struct LoadU32as64 {
    const uint32_t* in;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return (uint64_t)in[i]; }
};
struct LoadU64 {
    const uint64_t* in;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return in[i]; }
};
struct LoadSegMax {
    const uint32_t* key;
    const int32_t* val;
    __device__ __forceinline__ SegMaxT operator()(uint64_t i) const {
        SegMaxT r;
        r.v = val[i];
        r.f = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
        return r;
    }
};
struct StoreU32 {
    uint32_t* out;
    __device__ __forceinline__ void operator()(uint64_t i, uint32_t v) const { out[i] = v; }
};
struct StoreU64 {
    uint64_t* out;
    __device__ __forceinline__ void operator()(uint64_t i, uint64_t v) const { out[i] = v; }
};
struct StoreSegMax {
    int32_t* out;
    __device__ __forceinline__ void operator()(uint64_t i, SegMaxT v) const { out[i] = v.v; }
};

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
template <typename T, typename Op, typename Load>
__global__ void __launch_bounds__(SC_THREADS) tile_reduce_kernel(Load ld, uint64_t n, T* __restrict__ partials) {
    __shared__ T sm[SC_THREADS / 32];
    Op op;
    const uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
    T acc = Op::identity();
#pragma unroll
    for (int j = 0; j < SC_ITEMS; j++)
        if (base + j < n) acc = op(acc, ld(base + j));
    T tot;
    block_inclusive_scan<T, Op, SC_THREADS>(acc, op, sm, &tot);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

// one CTA: in-place exclusive scan of the tile aggregates; optional grand total
template <typename T, typename Op>
__global__ void __launch_bounds__(1024) partials_scan_kernel(T* __restrict__ partials, uint64_t nb, T* total_out) {
    __shared__ T sm[32];
    __shared__ T inc_sm[1024];
    Op op;
    T carry = Op::identity();
    for (uint64_t base = 0; base < nb; base += 1024) {
        uint64_t i = base + threadIdx.x;
        T x = i < nb ? partials[i] : Op::identity();
        T tot;
        T inc = block_inclusive_scan<T, Op, 1024>(x, op, sm, &tot);
        inc_sm[threadIdx.x] = inc;
        __syncthreads();
        T ex = threadIdx.x == 0 ? Op::identity() : inc_sm[threadIdx.x - 1];
        if (i < nb) partials[i] = op(carry, ex);
        carry = op(carry, tot);
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename T, typename Op, typename Load, typename Store, bool INCLUSIVE>
__global__ void __launch_bounds__(SC_THREADS) tile_scan_kernel(Load ld, Store stf, uint64_t n,
                                                               const T* __restrict__ partials) {
    __shared__ T sm[SC_THREADS / 32];
    Op op;
    const uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
    T x[SC_ITEMS];
    T acc = Op::identity();
#pragma unroll
    for (int j = 0; j < SC_ITEMS; j++) {
        x[j] = (base + j < n) ? ld(base + j) : Op::identity();
        acc = op(acc, x[j]);
    }
    T tot;
    T inc = block_inclusive_scan<T, Op, SC_THREADS>(acc, op, sm, &tot);
    // exclusive prefix of this thread = carry-in (+) inclusive of the previous thread
    __shared__ T prev_inc[SC_THREADS];
    prev_inc[threadIdx.x] = inc;
    __syncthreads();
    T run = partials[blockIdx.x];
    if (threadIdx.x > 0) run = op(run, prev_inc[threadIdx.x - 1]);
#pragma unroll
    for (int j = 0; j < SC_ITEMS; j++) {
        if (INCLUSIVE) {
            run = op(run, x[j]);
            if (base + j < n) stf(base + j, run);
        } else {
            if (base + j < n) stf(base + j, run);
            run = op(run, x[j]);
        }
    }
}

// Short inputs (the n / 256 chunk totals of the pair join, per-key tables): one CTA, one launch instead of three.  Every thread
// owns a contiguous slice, reduces it, the CTA scans the 1024 slice totals, every thread re-reads its slice and writes.
constexpr uint64_t SC_SMALL_MAX = 4096;  // (at 39063 elements the single CTA takes 30 us against 12 us for the three launches)
template <typename T, typename Op, typename Load, typename Store, bool INCLUSIVE>
__global__ void __launch_bounds__(1024) small_scan_kernel(Load ld, Store stf, uint64_t n, T* total_out) {
    __shared__ T sm[32];
    __shared__ T inc_sm[1024];
    Op op;
    const uint64_t per = (n + 1023) / 1024;
    const uint64_t b = min(n, (uint64_t)threadIdx.x * per), e = min(n, b + per);
    T acc = Op::identity();
    for (uint64_t i = b; i < e; i++) acc = op(acc, ld(i));
    T tot;
    const T inc = block_inclusive_scan<T, Op, 1024>(acc, op, sm, &tot);
    inc_sm[threadIdx.x] = inc;
    __syncthreads();
    T run = threadIdx.x ? inc_sm[threadIdx.x - 1] : Op::identity();
    for (uint64_t i = b; i < e; i++) {
        const T x = ld(i);
        if (INCLUSIVE) { run = op(run, x); stf(i, run); }
        else { stf(i, run); run = op(run, x); }
    }
    if (threadIdx.x == 0 && total_out) *total_out = tot;
}

template <typename T, typename Op, typename Load, typename Store, bool INCLUSIVE>
static int run_scan(Load ld, Store stf, uint64_t n, T* scratch, T* total_out, cudaStream_t st) {
    if (n == 0) {
        if (total_out) {
            BRG_CUDA(cudaMemsetAsync(total_out, 0, sizeof(T), st));
        }
        return 0;
    }
    if (n <= SC_SMALL_MAX) {
        small_scan_kernel<T, Op, Load, Store, INCLUSIVE><<<1, 1024, 0, st>>>(ld, stf, n, total_out);
        BRG_LAUNCH_CHECK();
        return 0;
    }
    const uint64_t nb = (n + SC_TILE - 1) / SC_TILE;
    tile_reduce_kernel<T, Op, Load><<<(unsigned)nb, SC_THREADS, 0, st>>>(ld, n, scratch);
    BRG_LAUNCH_CHECK();
    partials_scan_kernel<T, Op><<<1, 1024, 0, st>>>(scratch, nb, total_out);
    BRG_LAUNCH_CHECK();
    tile_scan_kernel<T, Op, Load, Store, INCLUSIVE><<<(unsigned)nb, SC_THREADS, 0, st>>>(ld, stf, n, scratch);
    BRG_LAUNCH_CHECK();
    return 0;
}

uint64_t scan_scratch_words(uint64_t n) {
    // in 32-bit words; sized for the widest element (SegMaxT / u64 = 2 words), +2 for totals
    return 2 * ((n + SC_TILE - 1) / SC_TILE + 2);
}

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* scratch, cudaStream_t st) {
    return run_scan<uint32_t, SumOp<uint32_t>, LoadU32, StoreU32, false>(LoadU32{in}, StoreU32{out}, n, scratch, nullptr,
                                                                       st);
}
static int exclusive_scan_u32_total(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* scratch,
                                    uint32_t* total_out, cudaStream_t st) {
    return run_scan<uint32_t, SumOp<uint32_t>, LoadU32, StoreU32, false>(LoadU32{in}, StoreU32{out}, n, scratch,
                                                                       total_out, st);
}
int exclusive_scan_u32_to_u64(const uint32_t* in, uint64_t* out, uint64_t n, uint64_t* scratch, cudaStream_t st) {
    // out[n] receives the grand total
    return run_scan<uint64_t, SumOp<uint64_t>, LoadU32as64, StoreU64, false>(LoadU32as64{in}, StoreU64{out}, n, scratch,
                                                                           out + n, st);
}
int exclusive_scan_u64(const uint64_t* in, uint64_t* out, uint64_t n, uint64_t* scratch, cudaStream_t st) {
    // out[n] receives the grand total (chunk totals of the pair join: 256 rows of < 2^32 matches each)
    return run_scan<uint64_t, SumOp<uint64_t>, LoadU64, StoreU64, false>(LoadU64{in}, StoreU64{out}, n, scratch, out + n, st);
}
int segmented_running_max(const uint32_t* key, const int32_t* val, int32_t* out, uint64_t n, int32_t* scratch,
                          cudaStream_t st) {
    return run_scan<SegMaxT, SegMaxOp, LoadSegMax, StoreSegMax, true>(LoadSegMax{key, val}, StoreSegMax{out}, n,
                                                                    reinterpret_cast<SegMaxT*>(scratch), nullptr, st);
}

// ---------------------------------------------------------------------------------------------------
// merge_intervals (interval_tree.rs:52-73) on start-sorted rows: a row opens a new merged interval iff it is
// the first of its key or its start is > the running max end of everything before it in the key
// (`interval.first <= current.last` merges, so touching intervals merge and [1,5],[6,10] do not).
// ---------------------------------------------------------------------------------------------------
__global__ void merge_flags_kernel(const uint32_t* __restrict__ key, const int32_t* __restrict__ s_start,
                                   const int32_t* __restrict__ pme, uint64_t n, uint32_t* __restrict__ flags) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || key[i] != key[i - 1] || s_start[i] > pme[i - 1]) ? 1u : 0u;
}
__global__ void merge_write_kernel(const uint32_t* __restrict__ key, const int32_t* __restrict__ s_start,
                                   const int32_t* __restrict__ pme, const uint32_t* __restrict__ flags,
                                   const uint32_t* __restrict__ midx, uint64_t n, int32_t* __restrict__ m_first,
                                   int32_t* __restrict__ m_last, uint32_t* __restrict__ m_key) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = flags[i];
    uint32_t mi = midx[i] + h - 1;  // merged interval this row belongs to
    if (h) {
        m_first[mi] = s_start[i];
        m_key[mi] = key[i];
    }
    if (i + 1 == n || flags[i + 1]) m_last[mi] = pme[i];
}
__global__ void merged_width_kernel(const int32_t* __restrict__ m_first, const int32_t* __restrict__ m_last,
                                    uint64_t m, uint32_t* __restrict__ w) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int32_t d = wsub(m_last[i], m_first[i]);  // interior term of get_coverage: max(1, last - first) (interval_tree.rs:148)
    w[i] = (uint32_t)(d < 1 ? 1 : d);
}

int merge_sorted_intervals(const uint32_t* s_key, const int32_t* s_start, const int32_t* pme, uint64_t n,
                           uint32_t* d_flags, uint32_t* d_midx, uint32_t* scan_scratch, int32_t* m_first,
                           int32_t* m_last, uint32_t* m_key, uint32_t* d_n_merged, cudaStream_t st) {
    if (n == 0) {
        BRG_CUDA(cudaMemsetAsync(d_n_merged, 0, 4, st));
        return 0;
    }
    unsigned g = (unsigned)((n + 255) / 256);
    merge_flags_kernel<<<g, 256, 0, st>>>(s_key, s_start, pme, n, d_flags);
    BRG_LAUNCH_CHECK();
    int rc = exclusive_scan_u32_total(d_flags, d_midx, n, scan_scratch, d_n_merged, st);
    if (rc) return rc;
    merge_write_kernel<<<g, 256, 0, st>>>(s_key, s_start, pme, d_flags, d_midx, n, m_first, m_last, m_key);
    BRG_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// merge UDTF (merge.rs:260-358) on rows sorted by (key, start, end): an interval joins the current run when
// start <= run_end + min_dist (strict: <), boundary with a saturating i64 add.  With start <= end on every row the run's
// end equals the per-key running max (a new run starts above everything before it), so runs are found in parallel:
// flag -> scan -> compact; n_intervals = distance between consecutive run heads.
// ---------------------------------------------------------------------------------------------------
__global__ void merge_dist_flags_kernel(const uint32_t* __restrict__ key, const int32_t* __restrict__ s_start,
                                        const int32_t* __restrict__ pme, uint64_t n, long long min_dist, int strict,
                                        uint32_t* __restrict__ flags) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool head = (i == 0 || key[i] != key[i - 1]);
    if (!head) {
        const long long ce = pme[i - 1];
        long long boundary;  // ce.saturating_add(min_dist)
        if (min_dist > 0 && ce > LLONG_MAX - min_dist) boundary = LLONG_MAX;
        else if (min_dist < 0 && ce < LLONG_MIN - min_dist) boundary = LLONG_MIN;
        else boundary = ce + min_dist;
        const long long s = s_start[i];
        head = !(strict ? s < boundary : s <= boundary);
    }
    flags[i] = head ? 1u : 0u;
}
__global__ void merge_heads_kernel(const uint32_t* __restrict__ key, const int32_t* __restrict__ s_start,
                                   const int32_t* __restrict__ pme, const uint32_t* __restrict__ flags,
                                   const uint32_t* __restrict__ midx, uint64_t n, uint32_t* __restrict__ m_key,
                                   int32_t* __restrict__ m_first, int32_t* __restrict__ m_last, uint32_t* __restrict__ m_row0) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t h = flags[i];
    const uint32_t mi = midx[i] + h - 1;
    if (h) { m_first[mi] = s_start[i]; m_key[mi] = key[i]; m_row0[mi] = (uint32_t)i; }
    if (i + 1 == n || flags[i + 1]) m_last[mi] = pme[i];
}
__global__ void merge_counts_kernel(const uint32_t* __restrict__ m_row0, uint64_t m, uint64_t n, uint32_t* __restrict__ m_count) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    m_count[j] = (uint32_t)((j + 1 < m ? m_row0[j + 1] : (uint32_t)n) - m_row0[j]);
}
int merge_runs(const uint32_t* s_key, const int32_t* s_start, const int32_t* pme, uint64_t n, long long min_dist, int strict,
               uint32_t* d_flags, uint32_t* d_midx, uint32_t* scan_scratch, uint32_t* m_key, int32_t* m_first, int32_t* m_last,
               uint32_t* m_row0, uint32_t* d_n_merged, cudaStream_t st) {
    if (n == 0) { BRG_CUDA(cudaMemsetAsync(d_n_merged, 0, 4, st)); return 0; }
    const unsigned g = (unsigned)((n + 255) / 256);
    merge_dist_flags_kernel<<<g, 256, 0, st>>>(s_key, s_start, pme, n, min_dist, strict, d_flags);
    BRG_LAUNCH_CHECK();
    int rc = exclusive_scan_u32_total(d_flags, d_midx, n, scan_scratch, d_n_merged, st);
    if (rc) return rc;
    merge_heads_kernel<<<g, 256, 0, st>>>(s_key, s_start, pme, d_flags, d_midx, n, m_key, m_first, m_last, m_row0);
    BRG_LAUNCH_CHECK();
    return 0;
}
// cluster UDTF (cluster.rs:560-720): every sorted row gets the id and the extent of its run
__global__ void cluster_rows_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ midx,
                                    const int32_t* __restrict__ m_first, const int32_t* __restrict__ m_last, uint64_t n,
                                    uint32_t* __restrict__ cluster, int32_t* __restrict__ c_start, int32_t* __restrict__ c_end) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t mi = midx[i] + flags[i] - 1;
    cluster[i] = mi;
    c_start[i] = m_first[mi];
    c_end[i] = m_last[mi];
}
int cluster_rows(const uint32_t* flags, const uint32_t* midx, const int32_t* m_first, const int32_t* m_last, uint64_t n,
                 uint32_t* cluster, int32_t* c_start, int32_t* c_end, cudaStream_t st) {
    if (n == 0) return 0;
    cluster_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(flags, midx, m_first, m_last, n, cluster, c_start, c_end);
    BRG_LAUNCH_CHECK();
    return 0;
}
int merge_counts(const uint32_t* m_row0, uint64_t m, uint64_t n, uint32_t* m_count, cudaStream_t st) {
    if (m == 0) return 0;
    merge_counts_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m_row0, m, n, m_count);
    BRG_LAUNCH_CHECK();
    return 0;
}

int merged_width_prefix(const int32_t* m_first, const int32_t* m_last, uint64_t m, uint32_t* widths_tmp, int32_t* m_pre,
                        uint32_t* scan_scratch, cudaStream_t st) {
    if (m == 0) {
        BRG_CUDA(cudaMemsetAsync(m_pre, 0, 4, st));
        return 0;
    }
    merged_width_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m_first, m_last, m, widths_tmp);
    BRG_LAUNCH_CHECK();
    // wrapping u32 sums == the reference's wrapping i32 accumulator (interval_tree.rs:146-151)
    return exclusive_scan_u32_total(widths_tmp, reinterpret_cast<uint32_t*>(m_pre), m, scan_scratch,
                                    reinterpret_cast<uint32_t*>(m_pre) + m, st);
}

// ---------------------------------------------------------------------------------------------------
// segment offsets: off[c] = first row whose key is >= c, for c = 0..n_keys  (keys >= n_keys: invalid tail)
// ---------------------------------------------------------------------------------------------------
__global__ void segment_offsets_kernel(const uint32_t* __restrict__ key, uint64_t n, uint32_t n_keys,
                                       uint32_t* __restrict__ off) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int64_t kprev = (i == 0) ? -1 : (int64_t)min(key[i - 1], n_keys);
    int64_t kcur = (i == n) ? (int64_t)n_keys : (int64_t)min(key[i], n_keys);
    for (int64_t c = kprev + 1; c <= kcur; c++) off[c] = (uint32_t)i;
}
__global__ void flag_inverted_kernel(const int32_t* __restrict__ s, const int32_t* __restrict__ e, uint64_t n,
                                     uint32_t* __restrict__ flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = i < n && s[i] > e[i];
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}
int flag_inverted_rows(const int32_t* start, const int32_t* end, uint64_t n, uint32_t* d_flag, cudaStream_t st) {
    BRG_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), st));
    if (n == 0) return 0;
    flag_inverted_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(start, end, n, d_flag);
    BRG_LAUNCH_CHECK();
    return 0;
}

__global__ void pack_records_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                                    const int32_t* __restrict__ c, const int32_t* __restrict__ d, uint64_t n,
                                    int4* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = make_int4(a[i], b[i], c[i], d ? d[i] : 0);
}
int pack_records(const int32_t* a, const int32_t* b, const int32_t* c, const int32_t* d, uint64_t n, int4* out,
                 cudaStream_t st) {
    if (n == 0) return 0;
    pack_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, b, c, d, n, out);
    BRG_LAUNCH_CHECK();
    return 0;
}

int segment_offsets(const uint32_t* sorted_key, uint64_t n, uint32_t n_keys, uint32_t* off, cudaStream_t st) {
    segment_offsets_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(sorted_key, n, n_keys, off);
    BRG_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// compact rank index (see brg_internal.cuh): one thread per 16-element block
// ---------------------------------------------------------------------------------------------------
__global__ void ci_build_kernel(const int32_t* __restrict__ A, const uint32_t* __restrict__ off,
                                const uint32_t* __restrict__ boff, uint32_t n_keys, uint32_t n_blocks,
                                int32_t* __restrict__ S, uint4* __restrict__ blocks) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_blocks) return;
    uint32_t lo = 0, hi = n_keys;  // key = last k with boff[k] <= g (keys without blocks share their successor's offset)
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (boff[mid + 1] <= g) lo = mid + 1; else hi = mid;
    }
    const uint32_t k = lo;
    const uint32_t e0 = off[k] + (g - boff[k]) * CI_BLOCK, e1 = min(e0 + CI_BLOCK, off[k + 1]);
    const int32_t base = A[e0];
    S[g] = base;
    uint32_t d[CI_BLOCK - 1];
    bool wide = false;
#pragma unroll
    for (int t = 1; t < CI_BLOCK; t++) {
        const uint32_t idx = e0 + t;
        if (idx < e1) {
            const uint32_t v = (uint32_t)A[idx] - (uint32_t)base;
            wide |= v > 0xFFFDu;
            d[t - 1] = v & 0xFFFFu;
        } else {
            d[t - 1] = 0xFFFEu;
        }
    }
    if (wide) d[0] = 0xFFFFu;
    blocks[g] = make_uint4((uint32_t)base, d[0] | (d[1] << 16), d[2] | (d[3] << 16), d[4] | (d[5] << 16));
}
int ci_build(const int32_t* A, const uint32_t* off, const uint32_t* boff, uint32_t n_keys, uint32_t n_blocks, int32_t* S,
             uint4* blocks, cudaStream_t st) {
    if (n_blocks == 0) return 0;
    ci_build_kernel<<<(n_blocks + 255) / 256, 256, 0, st>>>(A, off, boff, n_keys, n_blocks, S, blocks);
    BRG_LAUNCH_CHECK();
    return 0;
}
__global__ void ci_patch_keydir_kernel(KeyDir* __restrict__ kd, const uint32_t* __restrict__ off, uint32_t n_keys) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_keys) return;
    kd[c].pad[0] = off[c];
    kd[c].pad[1] = off[c + 1];
}
// reads D[i + 1] while other threads overwrite it: the packed value keeps the original in its upper 30 bits, but to stay
// free of races the kernel writes to a second buffer
__global__ void ci_pack_table_kernel(const uint32_t* __restrict__ D, uint32_t total, uint32_t* __restrict__ P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t a = D[i];
    const uint32_t b = (i + 1 < total) ? D[i + 1] : a;
    const uint32_t cnt = b >= a ? min(b - a, 3u) : 0u;  // (b < a only across a key boundary: that entry is never read)
    P[i] = (a << 2) | cnt;
}
int ci_pack_table(const uint32_t* D, uint32_t total_entries, uint32_t* P, cudaStream_t st) {
    if (total_entries == 0) return 0;
    ci_pack_table_kernel<<<(total_entries + 255) / 256, 256, 0, st>>>(D, total_entries, P);
    BRG_LAUNCH_CHECK();
    return 0;
}
int ci_patch_keydir(KeyDir* kd, const uint32_t* off, uint32_t n_keys, cudaStream_t st) {
    if (n_keys == 0) return 0;
    ci_patch_keydir_kernel<<<(n_keys + 255) / 256, 256, 0, st>>>(kd, off, n_keys);
    BRG_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// count bins (see CountBin in brg_internal.cuh)
// ---------------------------------------------------------------------------------------------------
__global__ void cb_key_ranges_kernel(const int32_t* __restrict__ S, const int32_t* __restrict__ E, const uint32_t* __restrict__ off,
                                     uint32_t n_keys, KeyDir* __restrict__ kd, unsigned long long* __restrict__ bins_by_shift) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_keys) return;
    uint32_t b = off[c], e = off[c + 1];
    KeyDir k;
    k.seg_begin = b; k.seg_end = e; k.base = 0; k.pad[0] = k.pad[1] = k.pad[2] = 0;
    if (b == e) {
        k.lo = 0; k.nb = 0;
    } else {
        // (starts <= ends row by row unless the table holds inverted rows: take both arrays into account either way)
        const int32_t lo = min(S[b], E[b]), hi = max(S[e - 1], E[e - 1]);
        const uint32_t range = (uint32_t)hi - (uint32_t)lo;
        k.lo = lo;
        k.nb = range;  // provisional: the range; dir_assign_bases turns it into a bin count
        for (int s = 0; s < 32; s++) atomicAdd(&bins_by_shift[s], (unsigned long long)(range >> s) + 1ull);
    }
    kd[c] = k;
}
int cb_key_ranges(const int32_t* s_start, const int32_t* e_end, const uint32_t* off, uint32_t n_keys, KeyDir* kd,
                  unsigned long long* d_bins, cudaStream_t st) {
    BRG_CUDA(cudaMemsetAsync(d_bins, 0, 32 * sizeof(unsigned long long), st));
    if (n_keys == 0) return 0;
    cb_key_ranges_kernel<<<(n_keys + 255) / 256, 256, 0, st>>>(s_start, e_end, off, n_keys, kd, d_bins);
    BRG_LAUNCH_CHECK();
    return 0;
}
// key that owns record g (keys without records share the base of their successor); returns false for padding
__device__ __forceinline__ bool cb_owner(const KeyDir* __restrict__ kd, uint32_t n_keys, uint32_t g, uint32_t* key) {
    uint32_t lo = 0, hi = n_keys;  // last key with base <= g
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (kd[mid].base <= g) lo = mid; else hi = mid;
    }
    uint32_t c = lo;
    while (kd[c].nb == 0 || g - kd[c].base > kd[c].nb) {
        if (c == 0) return false;
        c--;
    }
    *key = c;
    return true;
}
// pass 1, one thread per record: cs / ce = first start / end >= the bin's first coordinate, through the rank directories of
// the two arrays (3-4 loads each; a plain binary search over a 400k-row segment is 19 dependent loads: 0.40 ms per 6M records)
__global__ void cb_bounds_kernel(DirView ds, DirView de, const KeyDir* __restrict__ kd, uint32_t n_keys, uint32_t shift,
                                 uint32_t* __restrict__ cs, uint32_t* __restrict__ ce, uint32_t total) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    uint32_t c;
    if (!cb_owner(kd, n_keys, g, &c)) { cs[g] = 0; ce[g] = 0; return; }
    const KeyDir k = kd[c];
    const uint32_t j = g - k.base;
    if (j >= k.nb) { cs[g] = k.seg_end; ce[g] = k.seg_end; return; }  // the key's sentinel record
    const int64_t thr = (int64_t)k.lo + ((int64_t)j << shift);  // <= the key's largest value: fits i32
    cs[g] = rank_lt_kd(ds, load_keydir(ds.kd, c), (int32_t)thr);
    ce[g] = rank_lt_kd(de, load_keydir(de.kd, c), (int32_t)thr);
}
// pass 2: the records (the bin's elements are S[cs[g] .. cs[g + 1]) and E[ce[g] .. ce[g + 1])); layout and formats: brg_bins.cuh
__global__ void cb_fill_kernel(const int32_t* __restrict__ S, const int32_t* __restrict__ E, const KeyDir* __restrict__ kd,
                               uint32_t n_keys, uint32_t shift, uint32_t fmt, const uint32_t* __restrict__ csv,
                               const uint32_t* __restrict__ cev, uint4* __restrict__ rec, uint32_t total) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    uint32_t c;
    if (!cb_owner(kd, n_keys, g, &c)) return;
    const KeyDir k = kd[c];
    const uint32_t j = g - k.base;
    uint32_t w[8] = {k.seg_end, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (j < k.nb) {
        const int64_t thr = (int64_t)k.lo + ((int64_t)j << shift);
        const uint32_t cs = csv[g], ce = cev[g];
        const uint32_t cs1 = csv[g + 1], ce1 = cev[g + 1];  // (record g + 1 is the next bin of this key or its sentinel)
        const uint32_t ns = cs1 - cs, ne = ce1 - ce;
        // depth = cs - ce: intervals that start before the bin and have not ended (negative only with inverted rows: overflow)
        if (cb_encode_header(w, fmt, cs, ce, ns, ne)) {
            for (uint32_t i = 0; i < ns + ne; i++) {
                const int64_t v = (int64_t)(i < ns ? S[cs + i] : E[ce + (i - ns)]) - thr;  // 0 .. 2^shift - 1
                cb_encode_slot(w, fmt, i, (uint32_t)v);
            }
        }
    }
    rec[2ull * g] = make_uint4(w[0], w[1], w[2], w[3]);
    rec[2ull * g + 1] = make_uint4(w[4], w[5], w[6], w[7]);
}
int cb_fill(const DirView& ds, const DirView& de, const KeyDir* kd, uint32_t n_keys, uint32_t shift, uint32_t fmt, uint32_t* cs_tmp,
            uint32_t* ce_tmp, uint4* rec, uint32_t total, cudaStream_t st) {
    if (total == 0 || n_keys == 0) return 0;
    cb_bounds_kernel<<<(total + 255) / 256, 256, 0, st>>>(ds, de, kd, n_keys, shift, cs_tmp, ce_tmp, total);
    BRG_LAUNCH_CHECK();
    cb_fill_kernel<<<(total + 255) / 256, 256, 0, st>>>(ds.A, de.A, kd, n_keys, shift, fmt, cs_tmp, ce_tmp, rec, total);
    BRG_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// rank directory construction
// ---------------------------------------------------------------------------------------------------
__global__ void dir_key_ranges_kernel(const int32_t* __restrict__ A, const uint32_t* __restrict__ off, uint32_t n_keys,
                                      KeyDir* __restrict__ kd, unsigned long long* __restrict__ bins_by_shift) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_keys) return;
    uint32_t b = off[c], e = off[c + 1];
    KeyDir k;
    k.seg_begin = b; k.seg_end = e; k.base = 0; k.pad[0] = k.pad[1] = k.pad[2] = 0;
    if (b == e) {
        k.lo = 0; k.nb = 0;
    } else {
        int32_t lo = A[b], hi = A[e - 1];
        uint32_t range = (uint32_t)hi - (uint32_t)lo;
        k.lo = lo;
        k.nb = range;  // provisional: the range; dir_assign_bases turns it into a bin count
        for (int s = 0; s < 32; s++) atomicAdd(&bins_by_shift[s], (unsigned long long)(range >> s) + 1ull);
    }
    kd[c] = k;
}
int dir_key_ranges(const int32_t* A, const uint32_t* off, uint32_t n_keys, KeyDir* kd, unsigned long long* d_bins,
                   cudaStream_t st) {
    BRG_CUDA(cudaMemsetAsync(d_bins, 0, 32 * sizeof(unsigned long long), st));
    if (n_keys == 0) return 0;
    dir_key_ranges_kernel<<<(n_keys + 255) / 256, 256, 0, st>>>(A, off, n_keys, kd, d_bins);
    BRG_LAUNCH_CHECK();
    return 0;
}

// one CTA: nb = (range >> shift) + 1 for non-empty keys; base = exclusive prefix of (nb + 1)
__global__ void __launch_bounds__(1024) dir_assign_bases_kernel(KeyDir* __restrict__ kd, uint32_t n_keys, uint32_t shift,
                                                                uint32_t* __restrict__ total) {
    __shared__ uint32_t sm[32];
    SumOp<uint32_t> op;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_keys; base += 1024) {
        uint32_t c = base + threadIdx.x;
        uint32_t cnt = 0, nb = 0;
        bool live = c < n_keys && kd[c].seg_begin != kd[c].seg_end;
        if (live) {
            nb = (kd[c].nb >> shift) + 1;
            cnt = nb + 1;
        }
        uint32_t tot;
        uint32_t inc = block_inclusive_scan<uint32_t, SumOp<uint32_t>, 1024>(cnt, op, sm, &tot);
        if (c < n_keys) {
            kd[c].nb = nb;
            kd[c].base = carry + inc - cnt;
        }
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
int dir_assign_bases(KeyDir* kd, uint32_t n_keys, uint32_t shift, uint32_t* d_total, cudaStream_t st) {
    dir_assign_bases_kernel<<<1, 1024, 0, st>>>(kd, n_keys, shift, d_total);
    BRG_LAUNCH_CHECK();
    return 0;
}

// one thread per directory entry: binary-search its bin boundary inside the key's segment
__global__ void dir_fill_kernel(const int32_t* __restrict__ A, const KeyDir* __restrict__ kd, uint32_t n_keys,
                                uint32_t shift, uint32_t* __restrict__ D, uint32_t total) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    // last key with base <= g that owns entries
    uint32_t lo = 0, hi = n_keys;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (kd[mid].base <= g) lo = mid; else hi = mid;
    }
    // keys with no entries share the base of their successor: walk back to the owner
    uint32_t c = lo;
    while (kd[c].nb == 0 || g - kd[c].base > kd[c].nb) {
        if (c == 0) return;
        c--;
    }
    KeyDir k = kd[c];
    uint32_t j = g - k.base;
    if (j == k.nb) { D[g] = k.seg_end; return; }
    int64_t thr = (int64_t)k.lo + ((int64_t)j << shift);  // first value of bin j
    uint32_t a = k.seg_begin, b = k.seg_end;
    while (a < b) {
        uint32_t mid = (a + b) >> 1;
        if ((int64_t)A[mid] < thr) a = mid + 1; else b = mid;
    }
    D[g] = a;
}
int dir_fill(const int32_t* A, const KeyDir* kd, uint32_t n_keys, uint32_t shift, uint32_t* D, uint32_t total,
             cudaStream_t st) {
    if (total == 0 || n_keys == 0) return 0;
    dir_fill_kernel<<<(total + 255) / 256, 256, 0, st>>>(A, kd, n_keys, shift, D, total);
    BRG_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// complement UDTF (complement.rs:297-358): the gaps of a contig's merged input intervals inside each of its view
// intervals [vs, ve).  The reference walks, per view, the merged intervals with me > vs and ms < ve (a contiguous range
// [j0, j1) because ms and me are both non-decreasing), emits [cursor, max(ms, vs)) when that is non-empty, moves the
// cursor to min(me, ve), and finally emits [cursor, ve).  Inside the range a gap precedes merged interval j > j0 iff
// ms[j] > me[j-1] (always, unless `strict` left two touching runs unmerged), so with G = exclusive prefix sum of those
// flags a view's row count is  [ms[j0] > vs] + G[j1] - G[j0+1] + [cursor < ve]:  count -> scan -> emit, and the emit
// pass is parallel over OUTPUT rows (a chromosome-sized view holds millions of gaps): each output slot finds its view
// by a search in the view offsets and its gap by a search in G.
// Views whose key is absent from the input (view-only contigs, complement.rs:437-457) come out whole and unconditionally.
// ---------------------------------------------------------------------------------------------------
__global__ void gap_flags_kernel(const uint32_t* __restrict__ m_key, const int32_t* __restrict__ m_first,
                                 const int32_t* __restrict__ m_last, uint64_t m, uint32_t* __restrict__ g) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    g[j] = (j > 0 && m_key[j] == m_key[j - 1] && m_first[j] > m_last[j - 1]) ? 1u : 0u;
}
__global__ void complement_count_kernel(const uint32_t* __restrict__ v_key, const long long* __restrict__ v_start,
                                        const long long* __restrict__ v_end, uint64_t n_views,
                                        const uint32_t* __restrict__ m_off, uint32_t n_keys,
                                        const int32_t* __restrict__ m_first, const int32_t* __restrict__ m_last,
                                        const uint32_t* __restrict__ G, uint4* __restrict__ plan,
                                        uint32_t* __restrict__ counts) {
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_views) return;
    const uint32_t k = v_key[v];
    const long long vs = v_start[v], ve = v_end[v];
    uint32_t j0 = 0, j1 = 0, lead = 0, tail = 0, mid = 0;
    if (k >= n_keys || m_off[k] == m_off[k + 1]) {
        tail = 1;  // view-only contig (no input rows): (view_start, view_end) as is
    } else if (vs < ve) {
        uint32_t lo = m_off[k], hi = m_off[k + 1];
        const uint32_t seg_end = hi;
        while (lo < hi) {  // first merged interval with me > vs
            const uint32_t mid_j = (lo + hi) >> 1;
            if ((long long)m_last[mid_j] > vs) hi = mid_j; else lo = mid_j + 1;
        }
        j0 = lo;
        hi = seg_end;
        while (lo < hi) {  // first merged interval with ms >= ve
            const uint32_t mid_j = (lo + hi) >> 1;
            if ((long long)m_first[mid_j] >= ve) hi = mid_j; else lo = mid_j + 1;
        }
        j1 = lo;
        long long cursor = vs;
        if (j1 > j0) {
            lead = (long long)m_first[j0] > vs ? 1u : 0u;
            mid = G[j1] - G[j0 + 1];
            cursor = min((long long)m_last[j1 - 1], ve);
        }
        tail = cursor < ve ? 1u : 0u;
    }
    plan[v] = make_uint4(j0, j1, lead, tail);
    counts[v] = lead + mid + tail;
}
__global__ void complement_emit_kernel(const long long* __restrict__ v_start, const long long* __restrict__ v_end,
                                       uint64_t n_views, const uint64_t* __restrict__ offsets,
                                       const int32_t* __restrict__ m_first, const int32_t* __restrict__ m_last,
                                       const uint32_t* __restrict__ G, const uint4* __restrict__ plan, uint64_t total,
                                       uint32_t* __restrict__ out_view, long long* __restrict__ out_start,
                                       long long* __restrict__ out_end) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        uint64_t lo = 0, hi = n_views;  // last view whose first slot is <= t (views without rows share their successor's offset)
        while (hi - lo > 1) {
            const uint64_t mid_v = (lo + hi) >> 1;
            if (offsets[mid_v] <= t) lo = mid_v; else hi = mid_v;
        }
        const uint64_t v = lo;
        uint32_t r = (uint32_t)(t - offsets[v]);
        const uint4 p = plan[v];
        const long long vs = v_start[v], ve = v_end[v];
        long long s, e;
        if (p.z && r == 0) {
            s = vs; e = m_first[p.x];
        } else {
            r -= p.z;
            const uint32_t mid = p.y > p.x ? G[p.y] - G[p.x + 1] : 0u;
            if (r < mid) {
                const uint32_t target = G[p.x + 1] + r;  // the flagged j with exactly `r` flagged intervals before it in (j0, j1)
                uint32_t a = p.x + 1, b = p.y;           // = the last j with G[j] <= target
                while (b - a > 1) {
                    const uint32_t c = (a + b) >> 1;
                    if (G[c] <= target) a = c; else b = c;
                }
                s = m_last[a - 1]; e = m_first[a];
            } else {
                s = p.y > p.x ? min((long long)m_last[p.y - 1], ve) : vs;
                e = ve;
            }
        }
        out_view[t] = (uint32_t)v;
        out_start[t] = s;
        out_end[t] = e;
    }
}
int complement_plan(const uint32_t* m_key, const int32_t* m_first, const int32_t* m_last, uint64_t m, uint32_t n_keys,
                    const uint32_t* v_key, const long long* v_start, const long long* v_end, uint64_t n_views,
                    uint32_t* m_off /*n_keys+1*/, uint32_t* g_tmp /*m*/, uint32_t* G /*m+1*/, uint32_t* scan_scratch,
                    uint4* plan /*n_views*/, uint32_t* counts /*n_views*/, cudaStream_t st) {
    int rc = segment_offsets(m_key, m, n_keys, m_off, st);
    if (rc) return rc;
    if (m) {
        gap_flags_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m_key, m_first, m_last, m, g_tmp);
        BRG_LAUNCH_CHECK();
        if ((rc = exclusive_scan_u32_total(g_tmp, G, m, scan_scratch, G + m, st))) return rc;
    }
    if (n_views) {
        complement_count_kernel<<<(unsigned)((n_views + 255) / 256), 256, 0, st>>>(v_key, v_start, v_end, n_views, m_off, n_keys,
                                                                                  m_first, m_last, G, plan, counts);
        BRG_LAUNCH_CHECK();
    }
    return 0;
}
int complement_emit(const long long* v_start, const long long* v_end, uint64_t n_views, const uint64_t* offsets,
                    const int32_t* m_first, const int32_t* m_last, const uint32_t* G, const uint4* plan, uint64_t total,
                    uint32_t* out_view, long long* out_start, long long* out_end, cudaStream_t st) {
    if (total == 0) return 0;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((total + 255) / 256, (uint64_t)kSMs * 16));
    complement_emit_kernel<<<grid, 256, 0, st>>>(v_start, v_end, n_views, offsets, m_first, m_last, G, plan, total, out_view,
                                                 out_start, out_end);
    BRG_LAUNCH_CHECK();
    return 0;
}


}  // namespace brg
