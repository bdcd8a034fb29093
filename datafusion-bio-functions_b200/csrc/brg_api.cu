// brg_api.cu -- the C ABI of libbioranges.so (see include/bioranges.h) and the host-side orchestration:
// context / stream lanes, index build pipeline (K1 sort -> K2 scans -> rank directories -> one HBM slab),
// chunked host<->device pipelines for the probe entry points, variable-length results in pooled pinned memory.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <chrono>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/bioranges.h"
#include "brg_internal.cuh"
#include "brg_tile.cuh"

namespace brg {

static thread_local std::string t_error;
void set_error(const std::string& msg) { t_error = msg; }
std::atomic<uint64_t> g_launches{0};

// ---------------------------------------------------------------------------------------------------
// small RAII helpers (host code only)
// ---------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    uint64_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    int ensure(uint64_t b) {
        if (b <= bytes) return 0;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        if (b == 0) return 0;
        cudaError_t e = cudaMalloc(&p, b);
        if (e != cudaSuccess) { p = nullptr; set_error(std::string("cudaMalloc(") + std::to_string(b) + "): " + cudaGetErrorString(e)); cudaGetLastError(); return BRG_ERR_OOM; }
        bytes = b;
        return 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Temporaries of the build / sort + sweep pipelines: stream-ordered allocations from the device's default pool, whose release
// threshold is "keep" (brg_ctx_create), so after the first call an allocation is a pointer bump instead of a cudaMalloc +
// cudaFree pair (each of which costs 0.1 - 0.3 ms and synchronises the device; ~25 of them per build).
// A ScratchScope names the stream the buffers of the enclosing function live on.
static thread_local cudaStream_t tls_scratch_stream = nullptr;
struct ScratchScope {
    cudaStream_t prev;
    explicit ScratchScope(cudaStream_t st) : prev(tls_scratch_stream) { tls_scratch_stream = st; }
    ~ScratchScope() { tls_scratch_stream = prev; }
};
// The pool keeps what it is given back (that is what makes the second build fast), so the scratch of one very large build
// would stay reserved for good: every build starts by returning everything above 4 GB that earlier calls left unused.
static void trim_scratch_pool(int device) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolTrimTo(pool, 4ull << 30);
    cudaGetLastError();
}
struct ScratchBuf {
    void* p = nullptr;
    uint64_t bytes = 0;
    cudaStream_t st = nullptr;
    ~ScratchBuf() { if (p) cudaFreeAsync(p, st); }
    int ensure(uint64_t b) {
        if (b <= bytes) return 0;
        if (p) { cudaFreeAsync(p, st); p = nullptr; bytes = 0; }
        if (b == 0) return 0;
        st = tls_scratch_stream;
        cudaError_t e = cudaMallocAsync(&p, b, st);
        if (e != cudaSuccess) { p = nullptr; set_error(std::string("cudaMallocAsync(") + std::to_string(b) + "): " + cudaGetErrorString(e)); cudaGetLastError(); return BRG_ERR_OOM; }
        bytes = b;
        return 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBlock { void* p; uint64_t bytes; };

// page-locked host buffer owned by a lane (the staging ring of pageable caller buffers, key runs)
struct HostBuf {
    void* p = nullptr;
    uint64_t bytes = 0;
    ~HostBuf() { if (p) cudaFreeHost(p); }
    int ensure(uint64_t b) {
        if (b <= bytes) return 0;
        if (p) { cudaFreeHost(p); p = nullptr; bytes = 0; }
        cudaError_t e = cudaHostAlloc(&p, b, cudaHostAllocDefault);
        if (e != cudaSuccess) { p = nullptr; set_error(std::string("cudaHostAlloc(") + std::to_string(b) + "): " + cudaGetErrorString(e)); cudaGetLastError(); return BRG_ERR_OOM; }
        bytes = b;
        return 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Host worker threads of a context: the copies between pageable caller buffers and the pinned staging ring, and the scan of
// a probe key column for runs.  parallel_for(n, fn) runs fn(0..n-1) on the workers plus the calling thread and returns when
// all are done; one caller at a time (callers serialise on `call_mu`).
class HostPool {
public:
    explicit HostPool(int n_threads) {
        for (int i = 0; i < n_threads; i++) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int width() const { return (int)workers_.size() + 1; }
    void parallel_for(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        std::lock_guard<std::mutex> call(call_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; next_ = 0; total_ = n; pending_ = n; gen_++;
        }
        cv_.notify_all();
        run_some();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
private:
    void run_some() {
        for (;;) {
            int i;
            const std::function<void(int)>* f;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!fn_ || next_ >= total_) return;
                i = next_++;
                f = fn_;
            }
            (*f)(i);
            bool last;
            { std::lock_guard<std::mutex> lk(mu_); last = --pending_ == 0; }
            if (last) done_cv_.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            run_some();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int next_ = 0, total_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

struct Lane {
    cudaStream_t h2d = nullptr, k = nullptr, d2h = nullptr;
    static constexpr int NBUF = 2;
    DevBuf in_key[NBUF], in_s[NBUF], in_e[NBUF], out0[NBUF], out1[NBUF], out2[NBUF];
    cudaEvent_t ev_in[NBUF] = {}, ev_k[NBUF] = {}, ev_out[NBUF] = {};
    DevBuf scratch_a, scratch_b, scratch_c, scratch_d;  // counts / offsets / scan scratch / flags
    DevBuf raw_s[NBUF], raw_e[NBUF], bad_rows;          // Arrow ingest: position columns in their native width; first bad rows
    // pinned staging ring for pageable caller buffers (filled / drained by the context's host pool) and per-chunk key runs
    HostBuf stage_key[NBUF], stage_s[NBUF], stage_e[NBUF], stage_o0[NBUF], stage_o1[NBUF], stage_o2[NBUF], stage_runs[NBUF];
    DevBuf runs_key[NBUF], runs_end[NBUF];
    bool busy = false;
};

}  // namespace brg

using namespace brg;

struct brg_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t build_stream = nullptr;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Lane*> lanes;
    std::vector<PinnedBlock> pinned_free;
    uint64_t chunk_rows = 8ull << 20;
    // Variable-length results (brg_rows, brg_merged, brg_gaps) hand out pinned blocks of this context's pool and may outlive
    // brg_ctx_destroy (a garbage-collected host language releases them late): every result holds a reference; the struct is
    // freed by whoever drops the last one, and after `dead` blocks go straight back to the driver.
    std::atomic<int> refs{1};
    std::atomic<bool> dead{false};
    std::unique_ptr<HostPool> pool;  // created on first use
    int host_threads = 0;
};
static void ctx_unref(brg_ctx* c) {
    if (c && c->refs.fetch_sub(1) == 1) delete c;
}
static HostPool* host_pool(brg_ctx* c) {
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->pool) c->pool.reset(new HostPool(std::max(0, c->host_threads - 1)));
    return c->pool.get();
}

// serialisable description of the slab
struct DirDesc { uint64_t off_kd, off_D; uint32_t shift, entries; };
struct IndexHeader {
    uint32_t magic, version;
    uint32_t n_keys, n, n_merged, flags;
    uint64_t slab_bytes;
    uint64_t off_s_start, off_s_end, off_s_pos, off_pme, off_e_end, off_m_first, off_m_last, off_m_pre, off_s_rec, off_e_rec,
        off_m_rec;
    DirDesc dir[5];  // 0 s_start, 1 e_end, 2 pme, 3 m_first, 4 m_last
    // compact rank index of s_start / e_end (per-row count path): samples, delta blocks, directory over the samples
    uint64_t off_cS[2], off_cB[2];
    DirDesc cdir[2];
    uint32_t n_blocks, pad_;
    // count bins (32-byte records; cbdir.off_kd = per-key constants, cbdir.off_D = the records, cbdir.entries = record count)
    DirDesc cbdir;
    DirDesc cvdir;  // coverage bins over (m_first, m_last), same layout
    DirDesc nbdir;  // nearest / pair bins over (s_start, pme), same layout
    // (bins tables: DirDesc.shift = bin width (log2) | record format << 8, see brg_bins.cuh)
};
static const uint32_t kMagic = 0x31475242u;  // "BRG1"
static const uint32_t kVersion = 7;
static const uint64_t kNone = ~0ull;
static const uint64_t kSlabHeaderBytes = 4096;  // reserved at the front of every slab for the IndexHeader
static_assert(sizeof(IndexHeader) <= kSlabHeaderBytes, "IndexHeader must fit the reserved slab prefix");

struct brg_index {
    brg_ctx* ctx = nullptr;
    IndexHeader h{};
    void* slab = nullptr;
    bool pooled = false;  // slab came from the stream-ordered pool (brg_build): a second build reuses it without a cudaMalloc
    IndexView view{};
};

static IndexView make_view(const IndexHeader& h, void* slab) {
    char* b = reinterpret_cast<char*>(slab);
    auto P = [&](uint64_t off) -> void* { return off == kNone ? nullptr : (void*)(b + off); };
    IndexView v{};
    v.n_keys = h.n_keys; v.n = h.n; v.n_merged = h.n_merged; v.flags = h.flags;
    v.s_start = (const int32_t*)P(h.off_s_start); v.s_end = (const int32_t*)P(h.off_s_end);
    v.s_pos = (const uint32_t*)P(h.off_s_pos); v.pme = (const int32_t*)P(h.off_pme);
    v.e_end = (const int32_t*)P(h.off_e_end);
    v.s_rec = (const int4*)P(h.off_s_rec); v.e_rec = (const int4*)P(h.off_e_rec); v.m_rec = (const int4*)P(h.off_m_rec);
    v.m_first = (const int32_t*)P(h.off_m_first); v.m_last = (const int32_t*)P(h.off_m_last);
    v.m_pre = (const int32_t*)P(h.off_m_pre);
    const int32_t* arrs[5] = {v.s_start, v.e_end, v.pme, v.m_first, v.m_last};
    DirView* dv[5] = {&v.d_sstart, &v.d_eend, &v.d_pme, &v.d_mfirst, &v.d_mlast};
    for (int i = 0; i < 5; i++) {
        dv[i]->kd = (const KeyDir*)P(h.dir[i].off_kd);
        dv[i]->D = (const uint32_t*)P(h.dir[i].off_D);
        dv[i]->A = arrs[i];
        dv[i]->shift = h.dir[i].shift;
    }
    CompactView* cv[2] = {&v.c_sstart, &v.c_eend};
    const int32_t* plain[2] = {v.s_start, v.e_end};
    for (int a = 0; a < 2; a++) {
        cv[a]->blocks = (const uint4*)P(h.off_cB[a]);
        cv[a]->A = plain[a];
        cv[a]->d.kd = (const KeyDir*)P(h.cdir[a].off_kd);
        cv[a]->d.D = (const uint32_t*)P(h.cdir[a].off_D);
        cv[a]->d.A = (const int32_t*)P(h.off_cS[a]);
        cv[a]->d.shift = h.cdir[a].shift;
    }
    const DirDesc* bd[3] = {&h.cbdir, &h.cvdir, &h.nbdir};
    BinsView* bv[3] = {&v.cb, &v.cvb, &v.nb};
    for (int i = 0; i < 3; i++) {
        bv[i]->rec = (const uint4*)P(bd[i]->off_D);
        bv[i]->kd = (const KeyDir*)P(bd[i]->off_kd);
        bv[i]->shift = bd[i]->shift & 0xFFu;
        bv[i]->fmt = bd[i]->shift >> 8;
    }
    return v;
}

// ---------------------------------------------------------------------------------------------------
// small kernels used only by the orchestration
// ---------------------------------------------------------------------------------------------------
namespace brg {
__global__ void init_rows_kernel(const uint32_t* __restrict__ key, const int32_t* __restrict__ s,
                                 const int32_t* __restrict__ e, uint64_t n, uint32_t* k_out, uint32_t* s_out,
                                 uint32_t* e_out, uint32_t* pos_out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    k_out[i] = key ? key[i] : 0u;
    s_out[i] = (uint32_t)s[i];
    e_out[i] = (uint32_t)e[i];
    pos_out[i] = (uint32_t)i;
}
__global__ void pack_bits_kernel(const uint32_t* __restrict__ v, uint64_t n, uint32_t* __restrict__ bits) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t n_round = (n + 31) & ~31ull;
    if (i >= n_round) return;
    uint32_t b = __ballot_sync(0xFFFFFFFFu, i < n && v[i] != 0);
    if ((threadIdx.x & 31) == 0) bits[i >> 5] = b;
}
// key column of rows [row0, row0 + m) from run-length form: run r covers rows [run_end[r-1], run_end[r])
__global__ void expand_runs_kernel(const uint32_t* __restrict__ run_key, const unsigned long long* __restrict__ run_end,
                                   uint32_t n_runs, uint64_t row0, uint64_t m, uint32_t* __restrict__ key_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned long long row = row0 + i;
    uint32_t lo = 0, hi = n_runs;  // first run whose end is > row
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(run_end + mid) <= row) lo = mid + 1; else hi = mid;
    }
    key_out[i] = lo < n_runs ? __ldg(run_key + lo) : 0xFFFFFFFFu;
}
// Arrow ingest: PosArray::resolve (array_utils.rs:73-131) on the device.  A position column that crossed
// PCIe in its native width (Int64 / UInt32 / UInt64) is narrowed to the i32 the kernels read; the first row whose value
// does not fit (i32::try_from fails) is recorded with an atomicMin and reported by the host after the call.
// One 16-byte load per thread and step.
template <typename T> __device__ __forceinline__ bool fits_i32(T v);
template <> __device__ __forceinline__ bool fits_i32<long long>(long long v) { return v >= INT32_MIN && v <= INT32_MAX; }
template <> __device__ __forceinline__ bool fits_i32<unsigned int>(unsigned int v) { return v <= (unsigned int)INT32_MAX; }
template <> __device__ __forceinline__ bool fits_i32<unsigned long long>(unsigned long long v) { return v <= (unsigned long long)INT32_MAX; }
template <typename T>
__global__ void narrow_positions_kernel(const T* __restrict__ src, uint64_t m, uint64_t row0, int32_t* __restrict__ dst,
                                        unsigned long long* __restrict__ first_bad) {
    constexpr int V = 16 / sizeof(T);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_vec = m / V;
    unsigned long long bad = ~0ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(src) + i);
        T v[V];
        memcpy(v, &raw, 16);
        int32_t o[V];
#pragma unroll
        for (int j = 0; j < V; j++) {
            o[j] = (int32_t)v[j];
            if (!fits_i32<T>(v[j]) && bad == ~0ull) bad = row0 + i * V + j;
        }
        if (V == 2) *reinterpret_cast<int2*>(dst + i * V) = make_int2(o[0], o[V - 1]);
        else *reinterpret_cast<int4*>(dst + i * V) = make_int4(o[0], o[1 % V], o[2 % V], o[3 % V]);
    }
    const uint64_t t = n_vec * V + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // scalar tail
    if (t < m) {
        const T v = src[t];
        dst[t] = (int32_t)v;
        if (!fits_i32<T>(v) && row0 + t < bad) bad = row0 + t;
    }
    if (bad != ~0ull) atomicMin(first_bad, bad);
}
static int launch_narrow(const void* src, int pos_type, uint64_t m, uint64_t row0, int32_t* dst, unsigned long long* first_bad,
                         cudaStream_t st) {
    if (m == 0) return 0;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((m / 2 + 255) / 256 + 1, (uint64_t)kSMs * 16));
    switch (pos_type) {
        case BRG_POS_INT64: narrow_positions_kernel<long long><<<grid, 256, 0, st>>>((const long long*)src, m, row0, dst, first_bad); break;
        case BRG_POS_UINT32: narrow_positions_kernel<unsigned int><<<grid, 256, 0, st>>>((const unsigned int*)src, m, row0, dst, first_bad); break;
        case BRG_POS_UINT64: narrow_positions_kernel<unsigned long long><<<grid, 256, 0, st>>>((const unsigned long long*)src, m, row0, dst, first_bad); break;
        default: set_error("probe: unknown position column type"); return BRG_ERR_INVALID;
    }
    BRG_LAUNCH_CHECK();
    return 0;
}
// experiment switch BRG_NARROW_OUT=1: counts / coverage cross PCIe as 32-bit values and are widened by the host threads
__global__ void narrow_i64_kernel(const long long* __restrict__ in, uint64_t m, int32_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = (int32_t)in[i];
}
static inline uint64_t pos_width(int t) { return (t == BRG_POS_INT64 || t == BRG_POS_UINT64) ? 8 : 4; }
// largest m in [1, n] with offsets[m] <= max_rows (at least one probe row is always consumed)
__global__ void find_cut_kernel(const uint64_t* __restrict__ offsets, uint64_t n, uint64_t max_rows, uint64_t* out) {
    uint64_t lo = 1, hi = n;  // answer in [lo, hi]
    while (lo < hi) {
        uint64_t mid = (lo + hi + 1) >> 1;
        if (offsets[mid] <= max_rows) lo = mid; else hi = mid - 1;
    }
    out[0] = lo;
    out[1] = offsets[lo];
}
}  // namespace brg

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
extern "C" BRG_API const char* brg_last_error(void) { return t_error.c_str(); }
extern "C" BRG_API int brg_version(void) { return 100; }
extern "C" BRG_API uint64_t brg_kernel_launches(void) { return g_launches.load(); }

extern "C" BRG_API int brg_ctx_create(int device_id, brg_ctx** out) {
    if (!out) { set_error("brg_ctx_create: out is NULL"); return BRG_ERR_INVALID; }
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libbioranges has no CPU fallback)");
        cudaGetLastError();
        return BRG_ERR_CUDA;
    }
    if (device_id < 0 || device_id >= n_dev) { set_error("brg_ctx_create: bad device id"); return BRG_ERR_INVALID; }
    BRG_CUDA(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    BRG_CUDA(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major < 10) {
        set_error(std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                  "; libbioranges is built for sm_100a only");
        return BRG_ERR_CUDA;
    }
    std::unique_ptr<brg_ctx> c(new brg_ctx());  // (freed on every early return below)
    c->device = device_id;
    c->sms = prop.multiProcessorCount;
    BRG_CUDA(cudaStreamCreateWithFlags(&c->build_stream, cudaStreamNonBlocking));
    auto fail = [&](cudaError_t err, const char* what) {
        set_error(std::string(what) + ": " + cudaGetErrorString(err));
        cudaStreamDestroy(c->build_stream);
        return BRG_ERR_CUDA;
    };
    {   // stream-ordered scratch (cudaMallocAsync) must not go back to the OS between calls
        cudaMemPool_t pool;
        cudaError_t pe = cudaDeviceGetDefaultMemPool(&pool, device_id);
        if (pe != cudaSuccess) return fail(pe, "cudaDeviceGetDefaultMemPool");
        uint64_t keep = ~0ull;
        pe = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        if (pe != cudaSuccess) return fail(pe, "cudaMemPoolSetAttribute");
    }
    if (const char* env = getenv("BRG_CHUNK_ROWS")) {
        uint64_t v = strtoull(env, nullptr, 10);
        if (v >= 1024) c->chunk_rows = v;
    }
    {   // host worker threads (staging of pageable buffers, key-run scans): BRG_HOST_THREADS, default min(8, cores of the process)
        int t = 0;
        if (const char* env = getenv("BRG_HOST_THREADS")) t = atoi(env);
        if (t <= 0) {
            cpu_set_t set;
            int cores = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
            t = std::max(1, std::min(8, cores));
        }
        c->host_threads = std::min(t, 64);
    }
    *out = c.release();
    return BRG_OK;
}

static void lane_free(Lane* l) {
    for (int b = 0; b < Lane::NBUF; b++) {
        if (l->ev_in[b]) cudaEventDestroy(l->ev_in[b]);
        if (l->ev_k[b]) cudaEventDestroy(l->ev_k[b]);
        if (l->ev_out[b]) cudaEventDestroy(l->ev_out[b]);
    }
    if (l->h2d) cudaStreamDestroy(l->h2d);
    if (l->k) cudaStreamDestroy(l->k);
    if (l->d2h) cudaStreamDestroy(l->d2h);
    delete l;
}

extern "C" BRG_API void brg_ctx_destroy(brg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->dead.store(true);
        for (Lane* l : c->lanes) lane_free(l);
        c->lanes.clear();
        for (auto& b : c->pinned_free) cudaFreeHost(b.p);
        c->pinned_free.clear();
        c->pool.reset();
    }
    if (c->build_stream) { cudaStreamDestroy(c->build_stream); c->build_stream = nullptr; }
    ctx_unref(c);  // results still alive keep the (now empty) struct until they are released
}
extern "C" BRG_API int brg_ctx_device(const brg_ctx* c) { return c ? c->device : -1; }
extern "C" BRG_API int brg_ctx_sm_count(const brg_ctx* c) { return c ? c->sms : 0; }

extern "C" BRG_API int brg_pinned_alloc(uint64_t bytes, void** out) {
    if (!out) return BRG_ERR_INVALID;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { set_error(std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); cudaGetLastError(); *out = nullptr; return BRG_ERR_OOM; }
    return BRG_OK;
}
extern "C" BRG_API void brg_pinned_free(void* p) { if (p) cudaFreeHost(p); }

static int lane_acquire(brg_ctx* c, Lane** out) {
    std::unique_lock<std::mutex> lk(c->mu);
    for (;;) {
        for (Lane* l : c->lanes)
            if (!l->busy) { l->busy = true; *out = l; return 0; }
        if (c->lanes.size() < 4) {
            Lane* l = new Lane();
            cudaError_t e = cudaStreamCreateWithFlags(&l->h2d, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&l->k, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&l->d2h, cudaStreamNonBlocking);
            for (int b = 0; b < Lane::NBUF && e == cudaSuccess; b++) {
                e = cudaEventCreateWithFlags(&l->ev_in[b], cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l->ev_k[b], cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l->ev_out[b], cudaEventDisableTiming);
            }
            if (e != cudaSuccess) { set_error(std::string("lane create: ") + cudaGetErrorString(e)); lane_free(l); return BRG_ERR_CUDA; }
            l->busy = true;
            c->lanes.push_back(l);
            *out = l;
            return 0;
        }
        c->cv.wait(lk);
    }
}
static void lane_release(brg_ctx* c, Lane* l) {
    { std::lock_guard<std::mutex> lk(c->mu); l->busy = false; }
    c->cv.notify_one();
}
struct LaneGuard {
    brg_ctx* c; Lane* l = nullptr;
    explicit LaneGuard(brg_ctx* c_) : c(c_) {}
    int acquire() { return lane_acquire(c, &l); }
    ~LaneGuard() { if (l) lane_release(c, l); }
};

static int pinned_get(brg_ctx* c, uint64_t bytes, PinnedBlock* out) {
    if (bytes == 0) bytes = 1;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        int best = -1;
        for (int i = 0; i < (int)c->pinned_free.size(); i++)
            if (c->pinned_free[i].bytes >= bytes && (best < 0 || c->pinned_free[i].bytes < c->pinned_free[best].bytes)) best = i;
        if (best >= 0 && c->pinned_free[best].bytes <= 4 * bytes + (1u << 20)) {
            *out = c->pinned_free[best];
            c->pinned_free.erase(c->pinned_free.begin() + best);
            return 0;
        }
    }
    uint64_t cap = (bytes + (1u << 20) - 1) & ~((1ull << 20) - 1);
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, cap, cudaHostAllocDefault);
    if (e != cudaSuccess) { set_error(std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); cudaGetLastError(); return BRG_ERR_OOM; }
    out->p = p; out->bytes = cap;
    return 0;
}
static void pinned_put(brg_ctx* c, PinnedBlock b) {
    if (!b.p) return;
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->dead.load() || c->pinned_free.size() >= 16) { cudaFreeHost(b.p); return; }  // (dead: the context was destroyed first)
    c->pinned_free.push_back(b);
}

// ---------------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------------
static inline uint64_t align256(uint64_t x) { return (x + 255) & ~255ull; }

struct DirBuild {
    ScratchBuf kd;         // KeyDir[n_keys]
    uint32_t shift = 0;
    uint32_t entries = 0;
};

static const uint32_t kTargetPerBin = 4;

// phase 1 of a directory: key ranges + choice of the bin width
static int dir_plan(const int32_t* A, const uint32_t* off, uint32_t n_keys, uint64_t n_elems, DirBuild* db,
                    unsigned long long* d_bins, uint32_t* d_total, cudaStream_t st, uint64_t bins_wanted = 0) {
    int rc = db->kd.ensure((uint64_t)std::max(1u, n_keys) * sizeof(KeyDir));
    if (rc) return rc;
    rc = dir_key_ranges(A, off, n_keys, db->kd.as<KeyDir>(), d_bins, st);
    if (rc) return rc;
    unsigned long long bins[32];
    BRG_CUDA(cudaMemcpyAsync(bins, d_bins, sizeof(bins), cudaMemcpyDeviceToHost, st));
    BRG_CUDA(cudaStreamSynchronize(st));
    uint64_t budget = std::max<uint64_t>(bins_wanted ? bins_wanted : n_elems / kTargetPerBin, 2ull * n_keys);
    budget = std::max<uint64_t>(budget, 64);
    uint32_t shift = 31;
    for (uint32_t s = 0; s < 32; s++)
        if (bins[s] <= budget) { shift = s; break; }
    db->shift = shift;
    rc = dir_assign_bases(db->kd.as<KeyDir>(), n_keys, shift, d_total, st);
    if (rc) return rc;
    uint32_t total = 0;
    BRG_CUDA(cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, st));
    BRG_CUDA(cudaStreamSynchronize(st));
    db->entries = total;
    return 0;
}

extern "C" BRG_API int brg_build_dev(brg_ctx* ctx, const uint32_t* d_key, const int32_t* d_start, const int32_t* d_end,
                             uint64_t n, uint32_t n_keys, uint32_t want, void* stream, brg_index** out) {
    if (!ctx || !out) { set_error("brg_build: NULL ctx/out"); return BRG_ERR_INVALID; }
    *out = nullptr;
    if (n > 0 && (!d_start || !d_end)) { set_error("brg_build: NULL column"); return BRG_ERR_INVALID; }
    if ((want & BRG_WANT_ALL) == 0) { set_error("brg_build: empty want mask"); return BRG_ERR_INVALID; }
    if (n >= 0xFFFFFFFFull) {  // nearest.rs:357-361 / interval_join.rs:1624: row ids are UInt32
        set_error("left row index " + std::to_string(n - 1) + " exceeds UInt32 index capacity");
        return BRG_ERR_OVERFLOW;
    }
    if (n_keys == BRG_KEY_ABSENT) { set_error("brg_build: n_keys too large"); return BRG_ERR_INVALID; }
    BRG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->build_stream;
    want &= BRG_WANT_ALL;
    ScratchScope scratch_scope(st);
    trim_scratch_pool(ctx->device);
    // BRG_BUILD_TRACE=1: wall-clock time of the build stages on stderr (each mark synchronises the stream)
    static const bool trace = getenv("BRG_BUILD_TRACE") != nullptr;
    static const bool trace_sync = trace && getenv("BRG_BUILD_TRACE")[0] != '2';  // 2: host-side timestamps only
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace) return;
        if (trace_sync) cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[brg_build] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    const bool full = (want & (BRG_WANT_NEAREST | BRG_WANT_PAIRS)) != 0;
    const bool need_soa = (want & BRG_WANT_PAIRS) != 0;     // s_end / s_pos as plain arrays (pair walk)
    const bool need_rec = (want & BRG_WANT_NEAREST) != 0;   // s_rec / e_rec
    const bool need_e = (want & (BRG_WANT_COUNT | BRG_WANT_NEAREST | BRG_WANT_PAIRS)) != 0;  // pairs: count identity
    const bool need_pme = (want & (BRG_WANT_NEAREST | BRG_WANT_PAIRS | BRG_WANT_COVERAGE)) != 0;
    const bool need_cov = (want & BRG_WANT_COVERAGE) != 0;

    // ---- temporaries: two ping-pong row sets (key,start,end,pos) + scan/histogram scratch -----------
    const uint64_t nn = std::max<uint64_t>(n, 1);
    ScratchBuf setA[4], setB[4], hist, scan, small;
    int rc;
    for (int i = 0; i < 4; i++) {
        if ((rc = setA[i].ensure(nn * 4))) return rc;
        if ((rc = setB[i].ensure(nn * 4))) return rc;
    }
    if ((rc = hist.ensure(radix_hist_words(nn) * 4))) return rc;
    if ((rc = scan.ensure((scan_scratch_words(std::max<uint64_t>(radix_hist_words(nn), nn)) + 16) * 4))) return rc;
    if ((rc = small.ensure(4096))) return rc;
    unsigned long long* d_bins = small.as<unsigned long long>();           // 32 x u64
    uint32_t* d_scalar = reinterpret_cast<uint32_t*>(small.as<char>() + 512);  // misc u32 results

    mark("alloc temporaries");
    uint32_t* a[4] = {setA[0].as<uint32_t>(), setA[1].as<uint32_t>(), setA[2].as<uint32_t>(), setA[3].as<uint32_t>()};
    uint32_t* b[4] = {setB[0].as<uint32_t>(), setB[1].as<uint32_t>(), setB[2].as<uint32_t>(), setB[3].as<uint32_t>()};
    if (n) {
        init_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_key, d_start, d_end, n, a[0], a[1], a[2], a[3]);
        BRG_LAUNCH_CHECK();
    }
    // number of 8-bit passes that cover key ids 0..n_keys (n_keys itself = "invalid key" bucket)
    int key_passes = 0;
    for (uint64_t v = n_keys; v > 0; v >>= 8) key_passes++;

    // A count-only build keeps nothing but the two sorted value arrays (CountOverlapIndex discards the pairing,
    // interval_tree.rs:26-38): each sort then moves (key, value) -- two word arrays instead of four.
    const bool two_col = want == BRG_WANT_COUNT;
    ScratchBuf e_key, e_val, e_key2, e_val2;  // two_col: the (key, end) rows of the E-order sort and their ping-pong set
    if (two_col && n) {
        if ((rc = e_key.ensure(nn * 4)) || (rc = e_val.ensure(nn * 4)) || (rc = e_key2.ensure(nn * 4)) || (rc = e_val2.ensure(nn * 4))) return rc;
        BRG_CUDA(cudaMemcpyAsync(e_key.p, a[0], n * 4, cudaMemcpyDeviceToDevice, st));
        BRG_CUDA(cudaMemcpyAsync(e_val.p, a[2], n * 4, cudaMemcpyDeviceToDevice, st));
    }
    // ---- S-order: (key, start, end, pos) -----------------------------------------------------------------
    std::vector<SortPass> passes;
    if (full) for (int sh = 0; sh < 32; sh += 8) passes.push_back(SortPass{2, sh, 1, 0xFFFFFFFFu});
    for (int sh = 0; sh < 32; sh += 8) passes.push_back(SortPass{1, sh, 1, 0xFFFFFFFFu});
    for (int p = 0; p < key_passes; p++) passes.push_back(SortPass{0, 8 * p, 0, n_keys});
    int res = 0;
    rc = radix_sort_rows(a, b, two_col ? 2 : 4, n, passes.data(), (int)passes.size(), hist.as<uint32_t>(), hist.bytes / 4,
                         scan.as<uint32_t>(), st, &res);
    if (rc) return rc;
    uint32_t** S = res == 0 ? a : b;   // sorted rows  (two_col: only [0] key and [1] start are meaningful)
    uint32_t** T = res == 0 ? b : a;   // free set
    mark("S-order sort");

    // segment offsets per key; rows with key >= n_keys form the ignored tail
    ScratchBuf off, m_off;
    if ((rc = off.ensure(((uint64_t)n_keys + 2) * 4))) return rc;
    if ((rc = segment_offsets(S[0], n, n_keys, off.as<uint32_t>(), st))) return rc;
    uint32_t n_valid = 0;
    BRG_CUDA(cudaMemcpyAsync(&n_valid, off.as<uint32_t>() + n_keys, 4, cudaMemcpyDeviceToHost, st));
    BRG_CUDA(cudaStreamSynchronize(st));

    // pairs: does the identity #pairs = #(start <= qe) - #(end < qs) hold (no build row with start > end)?
    uint32_t idx_props = 0;
    if (need_soa && n_valid) {
        if ((rc = flag_inverted_rows((const int32_t*)S[1], (const int32_t*)S[2], n_valid, d_scalar + 8, st))) return rc;
        uint32_t inv = 0;
        BRG_CUDA(cudaMemcpyAsync(&inv, d_scalar + 8, 4, cudaMemcpyDeviceToHost, st));
        BRG_CUDA(cudaStreamSynchronize(st));
        if (inv) idx_props |= BRG_IDX_INVERTED;
    }

    // prefix_max_end over the S-order (per key)
    ScratchBuf pme;
    if (need_pme) {
        if ((rc = pme.ensure(nn * 4))) return rc;
        if ((rc = segmented_running_max(S[0], (const int32_t*)S[2], pme.as<int32_t>(), n_valid, scan.as<int32_t>(), st))) return rc;
    }

    // merged intervals (coverage)
    ScratchBuf m_first, m_last, m_key, m_pre, m_rec;
    uint32_t n_merged = 0;
    if (need_cov) {
        if ((rc = m_first.ensure(nn * 4)) || (rc = m_last.ensure(nn * 4)) || (rc = m_key.ensure(nn * 4)) ||
            (rc = m_pre.ensure((nn + 1) * 4)) || (rc = m_off.ensure(((uint64_t)n_keys + 2) * 4)))
            return rc;
        // T[1], T[2] are free: flags / merged index
        rc = merge_sorted_intervals(S[0], (const int32_t*)S[1], pme.as<int32_t>(), n_valid, T[1], T[2], scan.as<uint32_t>(),
                                    m_first.as<int32_t>(), m_last.as<int32_t>(), m_key.as<uint32_t>(), d_scalar, st);
        if (rc) return rc;
        BRG_CUDA(cudaMemcpyAsync(&n_merged, d_scalar, 4, cudaMemcpyDeviceToHost, st));
        BRG_CUDA(cudaStreamSynchronize(st));
        if ((rc = segment_offsets(m_key.as<uint32_t>(), n_merged, n_keys, m_off.as<uint32_t>(), st))) return rc;
        if ((rc = merged_width_prefix(m_first.as<int32_t>(), m_last.as<int32_t>(), n_merged, T[1], m_pre.as<int32_t>(),
                                      scan.as<uint32_t>(), st)))
            return rc;
        if ((rc = m_rec.ensure(std::max<uint64_t>(n_merged, 1) * 16))) return rc;
        if ((rc = pack_records(m_first.as<int32_t>(), m_last.as<int32_t>(), m_pre.as<int32_t>(), m_pre.as<int32_t>() + 1,
                               n_merged, m_rec.as<int4>(), st)))
            return rc;
    }

    mark("pme / merge");
    // keep the S-order columns: they are consumed by the E-order sort below
    ScratchBuf s_start, s_end, s_pos, s_rec, e_rec;
    if ((rc = s_start.ensure(nn * 4))) return rc;
    BRG_CUDA(cudaMemcpyAsync(s_start.p, S[1], (uint64_t)n_valid * 4, cudaMemcpyDeviceToDevice, st));
    if (need_soa) {
        if ((rc = s_end.ensure(nn * 4)) || (rc = s_pos.ensure(nn * 4))) return rc;
        BRG_CUDA(cudaMemcpyAsync(s_end.p, S[2], (uint64_t)n_valid * 4, cudaMemcpyDeviceToDevice, st));
        BRG_CUDA(cudaMemcpyAsync(s_pos.p, S[3], (uint64_t)n_valid * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (need_rec) {
        if ((rc = s_rec.ensure(nn * 16)) || (rc = e_rec.ensure(nn * 16))) return rc;
        if ((rc = pack_records((const int32_t*)S[1], (const int32_t*)S[2], (const int32_t*)S[3], pme.as<int32_t>(), n_valid,
                               s_rec.as<int4>(), st)))
            return rc;
    }

    // ---- E-order: (key, end, start, pos): stable sort of the S-order rows by end, then key -----------------
    uint32_t** E = nullptr;
    uint32_t* e2a[4] = {e_key.as<uint32_t>(), e_val.as<uint32_t>(), nullptr, nullptr};
    uint32_t* e2b[4] = {e_key2.as<uint32_t>(), e_val2.as<uint32_t>(), nullptr, nullptr};
    uint32_t* e2view[4] = {nullptr, nullptr, nullptr, nullptr};
    if (need_e && two_col) {
        // (key, end) of the ORIGINAL rows, sorted on their own: rows with an invalid key sort last, the first n_valid are kept
        passes.clear();
        for (int sh = 0; sh < 32; sh += 8) passes.push_back(SortPass{1, sh, 1, 0xFFFFFFFFu});
        for (int p = 0; p < key_passes; p++) passes.push_back(SortPass{0, 8 * p, 0, n_keys});
        int res2 = 0;
        rc = radix_sort_rows(e2a, e2b, 2, n, passes.data(), (int)passes.size(), hist.as<uint32_t>(), hist.bytes / 4,
                             scan.as<uint32_t>(), st, &res2);
        if (rc) return rc;
        uint32_t** R = res2 == 0 ? e2a : e2b;
        e2view[0] = R[0]; e2view[2] = R[1];  // E[2] = sorted ends, as in the four-column layout
        E = e2view;
    } else if (need_e) {
        passes.clear();
        for (int sh = 0; sh < 32; sh += 8) passes.push_back(SortPass{2, sh, 1, 0xFFFFFFFFu});
        for (int p = 0; p < key_passes; p++) passes.push_back(SortPass{0, 8 * p, 0, n_keys});
        int res2 = 0;
        rc = radix_sort_rows(S, T, 4, n_valid, passes.data(), (int)passes.size(), hist.as<uint32_t>(), hist.bytes / 4,
                             scan.as<uint32_t>(), st, &res2);
        if (rc) return rc;
        E = res2 == 0 ? S : T;
        if (need_rec && (rc = pack_records((const int32_t*)E[1], (const int32_t*)E[2], (const int32_t*)E[3], nullptr, n_valid,
                                           e_rec.as<int4>(), st)))
            return rc;
    }

    mark("E-order sort");
    // ---- rank directories ----------------------------------------------------------------------------------
    DirBuild db[5];
    const bool need_dir[5] = {(want & (BRG_WANT_COUNT | BRG_WANT_NEAREST | BRG_WANT_PAIRS)) != 0, need_e,
                              (want & (BRG_WANT_NEAREST | BRG_WANT_PAIRS)) != 0, need_cov, need_cov};
    const int32_t* dirA[5] = {s_start.as<int32_t>(), E ? (const int32_t*)E[2] : nullptr, pme.as<int32_t>(),
                              m_first.as<int32_t>(), m_last.as<int32_t>()};
    const uint32_t* dirOff[5] = {off.as<uint32_t>(), off.as<uint32_t>(), off.as<uint32_t>(), m_off.as<uint32_t>(),
                                 m_off.as<uint32_t>()};
    const uint64_t dirN[5] = {n_valid, n_valid, n_valid, n_merged, n_merged};
    for (int i = 0; i < 5; i++) {
        if (!need_dir[i]) continue;
        if ((rc = dir_plan(dirA[i], dirOff[i], n_keys, dirN[i], &db[i], d_bins, d_scalar, st))) return rc;
    }

    mark("directory plans");
    // ---- compact rank index of s_start / e_end: the L2-resident structure of the per-row count kernels -------
    // ---- count bins: the structure the per-row count kernels search (unsorted probe tables) ------------------------------
    // At most kCbTarget elements of each array per bin on average.  Offsets inside a bin are u16, so bins are at most 65536
    // wide: a build side too sparse for that (more than 2 records per row) keeps the compact rank index below instead.
    static const uint32_t kCbTarget = [] { const char* v = getenv("BRG_CB_TARGET"); const int t = v ? atoi(v) : 2; return (uint32_t)(t > 0 ? t : 0); }();
    // Record format (brg_bins.cuh): the packed one (18 slots of 10 bits, bins twice as wide for the same overflow odds) halves
    // a table; it pays when the wide-slot table would not stay in L2 under a shuffled probe table (10M build rows: 194 MB of
    // records -> 97 MB).  BRG_CB_PACK: 1 = whenever the bins are narrow enough, 0 = never, unset = tables above kPackAbove bytes.
    // (read at every build: the tests flip it inside one process)
    const int kCbPack = [] { const char* v = getenv("BRG_CB_PACK"); return v && v[0] ? atoi(v) : -1; }();
    static const uint64_t kPackAbove = 48ull << 20;
    // plan of one bins table over the array pair (A, B) that shares the key segments `seg`: per-key ranges, bin width, record count
    auto plan_bins = [&](const int32_t* A, const int32_t* B, const uint32_t* seg, uint64_t n_elems, ScratchBuf* kd, uint32_t* shift_out,
                         uint32_t* entries_out, bool* ok) -> int {
        int rc;
        *ok = false;
        if ((rc = kd->ensure((uint64_t)std::max(1u, n_keys) * sizeof(KeyDir)))) return rc;
        if ((rc = cb_key_ranges(A, B, seg, n_keys, kd->as<KeyDir>(), d_bins, st))) return rc;
        unsigned long long bins[32];
        BRG_CUDA(cudaMemcpyAsync(bins, d_bins, sizeof(bins), cudaMemcpyDeviceToHost, st));
        BRG_CUDA(cudaStreamSynchronize(st));
        // widest bins that still hold at most kCbTarget elements of each array on average (so between target / 2 and target)
        const uint64_t at_least = std::max<uint64_t>(n_elems / kCbTarget, 1);
        uint32_t shift = 0;
        for (uint32_t sft = CB_SHIFT_WIDE; sft > 0; sft--)
            if (bins[sft] >= at_least) { shift = sft; break; }
        if (bins[shift] > 2ull * n_elems + 2ull * n_keys + (1ull << 16)) return 0;  // (too sparse for 65536-wide bins)
        uint32_t fmt = 0;
        if (kCbPack != 0 && shift + 1 <= CB_SHIFT_PACKED && (kCbPack > 0 || bins[shift] * 32ull > kPackAbove)) { fmt = 1; shift += 1; }
        *shift_out = shift | (fmt << 8);
        if ((rc = dir_assign_bases(kd->as<KeyDir>(), n_keys, shift, d_scalar, st))) return rc;
        BRG_CUDA(cudaMemcpyAsync(entries_out, d_scalar, 4, cudaMemcpyDeviceToHost, st));
        BRG_CUDA(cudaStreamSynchronize(st));
        *ok = true;
        return 0;
    };
    ScratchBuf cb_kd, cv_kd, nb_kd;
    uint32_t cb_shift = 0, cb_entries = 0, cv_shift = 0, cv_entries = 0, nb_shift = 0, nb_entries = 0;  // (shift | format << 8)
    bool need_cb = (want & BRG_WANT_COUNT) != 0 && n_valid > 0 && kCbTarget > 0;
    if (need_cb && (rc = plan_bins(s_start.as<int32_t>(), (const int32_t*)E[2], off.as<uint32_t>(), n_valid, &cb_kd, &cb_shift, &cb_entries, &need_cb))) return rc;
    // coverage: the same bins over the merged intervals' (first, last) -- the per-row coverage kernels (unsorted probe tables)
    bool need_cvb = need_cov && n_merged > 0 && kCbTarget > 0;
    if (need_cvb && (rc = plan_bins(m_first.as<int32_t>(), m_last.as<int32_t>(), m_off.as<uint32_t>(), n_merged, &cv_kd, &cv_shift, &cv_entries, &need_cvb))) return rc;
    // nearest / pairs: bins over (s_start, prefix_max_end) -- the per-row kernels of those two (unsorted probe tables) read
    // #(start <= qe) and the first row whose prefix_max_end >= qs from one record.  Not with inverted build rows (start > end):
    // the kernels' shortcut "no overlap => every row that starts by qe has ended before qs" needs start <= end.
    // Built with BRG_WANT_NEAREST (50M x 5M shuffled nearest 3.39 -> 2.10 ms for +0.14 ms of build); a pairs-only index gets them
    // on request (BRG_NB_BINS=2): the shuffled 10M x 10M range pass goes 0.70 -> 0.38 ms, but the 97 MB larger slab made the warm
    // 10M-row pairs build 4.0 -> 6.4 ms (profiles/r2_bins_experiments.txt), which a join of sorted tables (0.2 ms) would pay for
    // nothing.  BRG_NB_BINS=0: never.  (read at every build: the tests flip it inside one process)
    const int kNbBins = [] { const char* v = getenv("BRG_NB_BINS"); return v && v[0] ? atoi(v) : 1; }();
    const uint32_t nb_wants = kNbBins >= 2 ? (BRG_WANT_NEAREST | BRG_WANT_PAIRS) : kNbBins == 1 ? BRG_WANT_NEAREST : 0u;
    bool need_nb = (want & nb_wants) != 0 && n_valid > 0 && kCbTarget > 0 && !(idx_props & BRG_IDX_INVERTED);
    if (need_nb && (rc = plan_bins(s_start.as<int32_t>(), pme.as<int32_t>(), off.as<uint32_t>(), n_valid, &nb_kd, &nb_shift, &nb_entries, &need_nb))) return rc;
    mark("count bins plan");
    const bool need_ci = (want & BRG_WANT_COUNT) != 0 && n_valid > 0 && !need_cb;
    ScratchBuf boff_d, cS[2], cB[2];
    DirBuild cdb[2];
    uint32_t n_blocks = 0;
    if (need_ci) {
        std::vector<uint32_t> h_off((size_t)n_keys + 1), h_boff((size_t)n_keys + 1);
        BRG_CUDA(cudaMemcpyAsync(h_off.data(), off.p, ((uint64_t)n_keys + 1) * 4, cudaMemcpyDeviceToHost, st));
        BRG_CUDA(cudaStreamSynchronize(st));
        uint64_t acc = 0;
        for (uint32_t k = 0; k < n_keys; k++) { h_boff[k] = (uint32_t)acc; acc += ((uint64_t)(h_off[k + 1] - h_off[k]) + CI_BLOCK - 1) / CI_BLOCK; }
        h_boff[n_keys] = (uint32_t)acc;
        n_blocks = (uint32_t)acc;
        if ((rc = boff_d.ensure(((uint64_t)n_keys + 2) * 4))) return rc;
        BRG_CUDA(cudaMemcpyAsync(boff_d.p, h_boff.data(), ((uint64_t)n_keys + 1) * 4, cudaMemcpyHostToDevice, st));
        for (int a = 0; a < 2; a++) {
            if ((rc = cS[a].ensure((uint64_t)n_blocks * 4)) || (rc = cB[a].ensure((uint64_t)n_blocks * 16))) return rc;
            if ((rc = ci_build(dirA[a], off.as<uint32_t>(), boff_d.as<uint32_t>(), n_keys, n_blocks, cS[a].as<int32_t>(), cB[a].as<uint4>(), st)))
                return rc;
            // up to 2 bins per block (the bin width is a power of two: 1..2 in practice)
            if ((rc = dir_plan(cS[a].as<int32_t>(), boff_d.as<uint32_t>(), n_keys, n_blocks, &cdb[a], d_bins, d_scalar, st, 2ull * n_blocks))) return rc;
            if ((rc = ci_patch_keydir(cdb[a].kd.as<KeyDir>(), off.as<uint32_t>(), n_keys, st))) return rc;
        }
        BRG_CUDA(cudaStreamSynchronize(st));  // h_boff is read by the copy above
    }

    mark("compact rank index");
    // ---- lay out and fill the slab -------------------------------------------------------------------------
    IndexHeader h{};
    h.magic = kMagic; h.version = kVersion;
    h.n_keys = n_keys; h.n = n_valid; h.n_merged = n_merged; h.flags = want | idx_props;
    uint64_t cur = kSlabHeaderBytes;  // the slab starts with a copy of its own header: one buffer describes and replicates itself
    auto place = [&](bool on, uint64_t bytes) -> uint64_t {
        if (!on) return kNone;
        uint64_t o = cur;
        cur = align256(cur + std::max<uint64_t>(bytes, 4));
        return o;
    };
    const uint64_t nv4 = (uint64_t)n_valid * 4;
    h.off_s_start = place(need_dir[0], nv4);
    h.off_s_end = place(need_soa, nv4);
    h.off_s_pos = place(need_soa, nv4);
    h.off_pme = place(need_dir[2], nv4);
    h.off_e_end = place(need_e, nv4);
    h.off_s_rec = place(need_rec, nv4 * 4);
    h.off_e_rec = place(need_rec, nv4 * 4);
    h.off_m_rec = place(need_cov, (uint64_t)n_merged * 16);
    h.off_m_first = place(need_cov, (uint64_t)n_merged * 4);
    h.off_m_last = place(need_cov, (uint64_t)n_merged * 4);
    h.off_m_pre = place(need_cov, ((uint64_t)n_merged + 1) * 4);
    for (int i = 0; i < 5; i++) {
        h.dir[i].off_kd = place(need_dir[i], (uint64_t)std::max(1u, n_keys) * sizeof(KeyDir));
        h.dir[i].off_D = place(need_dir[i], ((uint64_t)db[i].entries + 1) * 4);
        h.dir[i].shift = db[i].shift;
        h.dir[i].entries = db[i].entries;
    }
    h.n_blocks = n_blocks;
    for (int a = 0; a < 2; a++) {
        h.off_cS[a] = kNone;
        h.off_cB[a] = place(need_ci, (uint64_t)n_blocks * 16);
        h.cdir[a].off_kd = place(need_ci, (uint64_t)std::max(1u, n_keys) * sizeof(KeyDir));
        h.cdir[a].off_D = place(need_ci, ((uint64_t)cdb[a].entries + 1) * 4);
        h.cdir[a].shift = cdb[a].shift;
        h.cdir[a].entries = cdb[a].entries;
    }
    h.cbdir.off_kd = place(need_cb, (uint64_t)std::max(1u, n_keys) * sizeof(KeyDir));
    h.cbdir.off_D = place(need_cb, (uint64_t)cb_entries * 32);
    h.cbdir.shift = cb_shift;
    h.cbdir.entries = cb_entries;
    h.cvdir.off_kd = place(need_cvb, (uint64_t)std::max(1u, n_keys) * sizeof(KeyDir));
    h.cvdir.off_D = place(need_cvb, (uint64_t)cv_entries * 32);
    h.cvdir.shift = cv_shift;
    h.cvdir.entries = cv_entries;
    h.nbdir.off_kd = place(need_nb, (uint64_t)std::max(1u, n_keys) * sizeof(KeyDir));
    h.nbdir.off_D = place(need_nb, (uint64_t)nb_entries * 32);
    h.nbdir.shift = nb_shift;
    h.nbdir.entries = nb_entries;
    h.slab_bytes = std::max<uint64_t>(cur, 256);

    // The slab comes from the stream-ordered pool like the temporaries: cudaMalloc of a few hundred MB costs 1-3 ms (a third of
    // a warm 10M-row build), a pool allocation that reuses the slab of an index destroyed earlier costs microseconds.
    void* slab = nullptr;
    {
        cudaError_t e = cudaMallocAsync(&slab, h.slab_bytes, st);
        if (e != cudaSuccess) { set_error(std::string("cudaMallocAsync(index slab ") + std::to_string(h.slab_bytes) + "): " + cudaGetErrorString(e)); cudaGetLastError(); return BRG_ERR_OOM; }
    }
    auto slab_free = [&]() { cudaFreeAsync(slab, st); };
    char* base = reinterpret_cast<char*>(slab);
    auto copy_in = [&](uint64_t off_, const void* src, uint64_t bytes) -> cudaError_t {
        if (off_ == kNone || bytes == 0) return cudaSuccess;
        return cudaMemcpyAsync(base + off_, src, bytes, cudaMemcpyDeviceToDevice, st);
    };
    cudaError_t ce = cudaSuccess;
    auto chk = [&](cudaError_t e) { if (ce == cudaSuccess) ce = e; };
    chk(copy_in(h.off_s_start, s_start.p, nv4));
    chk(copy_in(h.off_s_end, s_end.p, nv4));
    chk(copy_in(h.off_s_pos, s_pos.p, nv4));
    chk(copy_in(h.off_pme, pme.p, nv4));
    if (E) chk(copy_in(h.off_e_end, E[2], nv4));
    chk(copy_in(h.off_s_rec, s_rec.p, nv4 * 4));
    chk(copy_in(h.off_e_rec, e_rec.p, nv4 * 4));
    chk(copy_in(h.off_m_rec, m_rec.p, (uint64_t)n_merged * 16));
    chk(copy_in(h.off_m_first, m_first.p, (uint64_t)n_merged * 4));
    chk(copy_in(h.off_m_last, m_last.p, (uint64_t)n_merged * 4));
    chk(copy_in(h.off_m_pre, m_pre.p, ((uint64_t)n_merged + 1) * 4));
    for (int i = 0; i < 5; i++)
        if (need_dir[i]) chk(copy_in(h.dir[i].off_kd, db[i].kd.p, (uint64_t)n_keys * sizeof(KeyDir)));
    if (need_ci)
        for (int a = 0; a < 2; a++) {
            chk(copy_in(h.off_cB[a], cB[a].p, (uint64_t)n_blocks * 16));
            chk(copy_in(h.cdir[a].off_kd, cdb[a].kd.p, (uint64_t)n_keys * sizeof(KeyDir)));
        }
    if (need_cb) chk(copy_in(h.cbdir.off_kd, cb_kd.p, (uint64_t)n_keys * sizeof(KeyDir)));
    if (need_cvb) chk(copy_in(h.cvdir.off_kd, cv_kd.p, (uint64_t)n_keys * sizeof(KeyDir)));
    if (need_nb) chk(copy_in(h.nbdir.off_kd, nb_kd.p, (uint64_t)n_keys * sizeof(KeyDir)));
    if (ce == cudaSuccess) ce = cudaMemsetAsync(base, 0, kSlabHeaderBytes, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(base, &h, sizeof(h), cudaMemcpyHostToDevice, st);  // (synchronised below, before h dies)
    if (ce != cudaSuccess) { slab_free(); set_error(std::string("slab copy: ") + cudaGetErrorString(ce)); return BRG_ERR_CUDA; }
    for (int i = 0; i < 5; i++) {
        if (!need_dir[i]) continue;
        rc = dir_fill(dirA[i], db[i].kd.as<KeyDir>(), n_keys, db[i].shift, (uint32_t*)(base + h.dir[i].off_D), db[i].entries, st);
        if (rc) { slab_free(); return rc; }
    }
    if (need_cb) {  // (after the rank directories of s_start and e_end: the record bounds are looked up through them)
        ScratchBuf cs_tmp, ce_tmp;
        if ((rc = cs_tmp.ensure(((uint64_t)cb_entries + 1) * 4)) || (rc = ce_tmp.ensure(((uint64_t)cb_entries + 1) * 4))) { slab_free(); return rc; }
        DirView ds, de;
        ds.kd = db[0].kd.as<KeyDir>(); ds.D = (const uint32_t*)(base + h.dir[0].off_D); ds.A = s_start.as<int32_t>(); ds.shift = db[0].shift;
        de.kd = db[1].kd.as<KeyDir>(); de.D = (const uint32_t*)(base + h.dir[1].off_D); de.A = (const int32_t*)E[2]; de.shift = db[1].shift;
        rc = cb_fill(ds, de, cb_kd.as<KeyDir>(), n_keys, cb_shift & 0xFFu, cb_shift >> 8, cs_tmp.as<uint32_t>(), ce_tmp.as<uint32_t>(),
                     (uint4*)(base + h.cbdir.off_D), cb_entries, st);
        if (rc) { slab_free(); return rc; }
    }
    if (need_nb) {  // (after the rank directories of s_start and pme)
        ScratchBuf cs_tmp, ce_tmp;
        if ((rc = cs_tmp.ensure(((uint64_t)nb_entries + 1) * 4)) || (rc = ce_tmp.ensure(((uint64_t)nb_entries + 1) * 4))) { slab_free(); return rc; }
        DirView ds, dp;
        ds.kd = db[0].kd.as<KeyDir>(); ds.D = (const uint32_t*)(base + h.dir[0].off_D); ds.A = s_start.as<int32_t>(); ds.shift = db[0].shift;
        dp.kd = db[2].kd.as<KeyDir>(); dp.D = (const uint32_t*)(base + h.dir[2].off_D); dp.A = pme.as<int32_t>(); dp.shift = db[2].shift;
        rc = cb_fill(ds, dp, nb_kd.as<KeyDir>(), n_keys, nb_shift & 0xFFu, nb_shift >> 8, cs_tmp.as<uint32_t>(), ce_tmp.as<uint32_t>(),
                     (uint4*)(base + h.nbdir.off_D), nb_entries, st);
        if (rc) { slab_free(); return rc; }
    }
    if (need_cvb) {
        ScratchBuf cs_tmp, ce_tmp;
        if ((rc = cs_tmp.ensure(((uint64_t)cv_entries + 1) * 4)) || (rc = ce_tmp.ensure(((uint64_t)cv_entries + 1) * 4))) { slab_free(); return rc; }
        DirView df, dl;
        df.kd = db[3].kd.as<KeyDir>(); df.D = (const uint32_t*)(base + h.dir[3].off_D); df.A = m_first.as<int32_t>(); df.shift = db[3].shift;
        dl.kd = db[4].kd.as<KeyDir>(); dl.D = (const uint32_t*)(base + h.dir[4].off_D); dl.A = m_last.as<int32_t>(); dl.shift = db[4].shift;
        rc = cb_fill(df, dl, cv_kd.as<KeyDir>(), n_keys, cv_shift & 0xFFu, cv_shift >> 8, cs_tmp.as<uint32_t>(), ce_tmp.as<uint32_t>(),
                     (uint4*)(base + h.cvdir.off_D), cv_entries, st);
        if (rc) { slab_free(); return rc; }
    }
    if (need_ci)
        for (int a = 0; a < 2; a++) {
            ScratchBuf dtmp;  // plain table first, then packed into the slab
            if ((rc = dtmp.ensure(((uint64_t)cdb[a].entries + 1) * 4))) { slab_free(); return rc; }
            rc = dir_fill(cS[a].as<int32_t>(), cdb[a].kd.as<KeyDir>(), n_keys, cdb[a].shift, dtmp.as<uint32_t>(), cdb[a].entries, st);
            if (!rc) rc = ci_pack_table(dtmp.as<uint32_t>(), cdb[a].entries, (uint32_t*)(base + h.cdir[a].off_D), st);
            if (!rc && cudaStreamSynchronize(st) != cudaSuccess) { set_error("compact table: sync failed"); rc = BRG_ERR_CUDA; }
            if (rc) { slab_free(); return rc; }
        }
    {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { slab_free(); set_error(std::string("build sync: ") + cudaGetErrorString(e)); return BRG_ERR_CUDA; }
    }
    mark("slab fill");
    brg_index* ix = new brg_index();
    ix->ctx = ctx; ix->h = h; ix->slab = slab; ix->pooled = true; ix->view = make_view(h, slab);
    *out = ix;
    return BRG_OK;
}

extern "C" BRG_API int brg_build(brg_ctx* ctx, const uint32_t* key, const int32_t* start, const int32_t* end, uint64_t n,
                         uint32_t n_keys, uint32_t want, brg_index** out) {
    if (!ctx || !out) { set_error("brg_build: NULL ctx/out"); return BRG_ERR_INVALID; }
    *out = nullptr;
    if (n > 0 && (!start || !end)) { set_error("brg_build: NULL column"); return BRG_ERR_INVALID; }
    BRG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->build_stream;
    ScratchScope scratch_scope(st);
    ScratchBuf dk, ds, de;
    int rc;
    const uint64_t bytes = std::max<uint64_t>(n, 1) * 4;
    if ((rc = ds.ensure(bytes)) || (rc = de.ensure(bytes))) return rc;
    if (key && (rc = dk.ensure(bytes))) return rc;
    if (n) {
        if (key) BRG_CUDA(cudaMemcpyAsync(dk.p, key, n * 4, cudaMemcpyHostToDevice, st));
        BRG_CUDA(cudaMemcpyAsync(ds.p, start, n * 4, cudaMemcpyHostToDevice, st));
        BRG_CUDA(cudaMemcpyAsync(de.p, end, n * 4, cudaMemcpyHostToDevice, st));
    }
    rc = brg_build_dev(ctx, key ? dk.as<uint32_t>() : nullptr, ds.as<int32_t>(), de.as<int32_t>(), n, n_keys, want, st, out);
    cudaStreamSynchronize(st);
    return rc;
}

extern "C" BRG_API void brg_index_destroy(brg_index* ix) {
    if (!ix) return;
    if (ix->ctx) cudaSetDevice(ix->ctx->device);
    if (ix->slab) {
        if (ix->pooled && ix->ctx && !ix->ctx->dead.load() && ix->ctx->build_stream) {
            cudaDeviceSynchronize();  // (what cudaFree does implicitly: probes still in flight on other streams finish first)
            cudaFreeAsync(ix->slab, ix->ctx->build_stream);
        } else {
            cudaFree(ix->slab);  // (also valid for pool memory)
        }
    }
    delete ix;
}
extern "C" BRG_API uint64_t brg_index_rows(const brg_index* ix) { return ix ? ix->h.n : 0; }
extern "C" BRG_API uint64_t brg_index_merged_rows(const brg_index* ix) { return ix ? ix->h.n_merged : 0; }
extern "C" BRG_API uint64_t brg_index_device_bytes(const brg_index* ix) { return ix ? ix->h.slab_bytes : 0; }
extern "C" BRG_API uint64_t brg_index_bins_info(const brg_index* ix, int table, uint32_t* bin_shift, uint32_t* format) {
    if (bin_shift) *bin_shift = 0;
    if (format) *format = 0;
    if (!ix || table < 0 || table > 2) return 0;
    const DirDesc& d = table == 0 ? ix->h.cbdir : table == 1 ? ix->h.cvdir : ix->h.nbdir;
    if (d.off_D == kNone) return 0;
    if (bin_shift) *bin_shift = d.shift & 0xFFu;
    if (format) *format = d.shift >> 8;
    return d.entries;
}

extern "C" BRG_API int brg_index_export(const brg_index* ix, void* header_buf, uint64_t cap, uint64_t* header_len, void** d_slab,
                                uint64_t* slab_bytes) {
    if (!ix) { set_error("brg_index_export: NULL index"); return BRG_ERR_INVALID; }
    if (header_len) *header_len = sizeof(IndexHeader);
    if (d_slab) *d_slab = ix->slab;
    if (slab_bytes) *slab_bytes = ix->h.slab_bytes;
    if (header_buf) {
        if (cap < sizeof(IndexHeader)) { set_error("brg_index_export: header buffer too small"); return BRG_ERR_INVALID; }
        memcpy(header_buf, &ix->h, sizeof(IndexHeader));
    }
    return BRG_OK;
}
extern "C" BRG_API int brg_index_import(brg_ctx* ctx, const void* header, uint64_t header_len, brg_index** out, void** d_slab,
                                uint64_t* slab_bytes) {
    if (!ctx || !header || !out) { set_error("brg_index_import: NULL argument"); return BRG_ERR_INVALID; }
    *out = nullptr;
    if (header_len != sizeof(IndexHeader)) { set_error("brg_index_import: header size mismatch"); return BRG_ERR_INVALID; }
    IndexHeader h;
    memcpy(&h, header, sizeof(h));
    if (h.magic != kMagic || h.version != kVersion) { set_error("brg_index_import: bad magic/version"); return BRG_ERR_INVALID; }
    BRG_CUDA(cudaSetDevice(ctx->device));
    void* slab = nullptr;
    cudaError_t e = cudaMalloc(&slab, h.slab_bytes);
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc(index slab): ") + cudaGetErrorString(e)); cudaGetLastError(); return BRG_ERR_OOM; }
    brg_index* ix = new brg_index();
    ix->ctx = ctx; ix->h = h; ix->slab = slab; ix->view = make_view(h, slab);
    *out = ix;
    if (d_slab) *d_slab = slab;
    if (slab_bytes) *slab_bytes = h.slab_bytes;
    return BRG_OK;
}

// One-collective replication (one process per GPU): the slab carries its header, so a receiver needs only the byte count up
// front: brg_index_slab_alloc -> receive the bytes (ncclBroadcast / cudaMemcpyPeer) -> brg_index_attach.
extern "C" BRG_API int brg_index_slab_alloc(brg_ctx* ctx, uint64_t slab_bytes, void** d_slab) {
    if (!ctx || !d_slab || slab_bytes < kSlabHeaderBytes) { set_error("brg_index_slab_alloc: bad argument"); return BRG_ERR_INVALID; }
    *d_slab = nullptr;
    BRG_CUDA(cudaSetDevice(ctx->device));
    // from the stream-ordered pool, like the slab of a build: a replica that replaces an earlier one of the same size costs
    // microseconds instead of a cudaMalloc of a few hundred MB (2-3 ms, more than the NVLink transfer itself)
    cudaError_t e = cudaMallocAsync(d_slab, slab_bytes, ctx->build_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->build_stream);  // usable on any stream from here on
    if (e != cudaSuccess) { set_error(std::string("cudaMallocAsync(index slab): ") + cudaGetErrorString(e)); cudaGetLastError(); *d_slab = nullptr; return BRG_ERR_OOM; }
    return BRG_OK;
}
extern "C" BRG_API void brg_index_slab_free(brg_ctx* ctx, void* d_slab) {
    if (!d_slab) return;
    if (ctx) cudaSetDevice(ctx->device);
    cudaFree(d_slab);  // (valid for pool memory as well)
}
extern "C" BRG_API int brg_index_attach(brg_ctx* ctx, void* d_slab, uint64_t slab_bytes, void* stream, brg_index** out) {
    if (!ctx || !d_slab || !out) { set_error("brg_index_attach: NULL argument"); return BRG_ERR_INVALID; }
    *out = nullptr;
    BRG_CUDA(cudaSetDevice(ctx->device));
    IndexHeader h;
    cudaStream_t st = (cudaStream_t)stream;
    BRG_CUDA(cudaMemcpyAsync(&h, d_slab, sizeof(h), cudaMemcpyDeviceToHost, st));  // ordered after the receive on `stream`
    BRG_CUDA(cudaStreamSynchronize(st));
    if (h.magic != kMagic || h.version != kVersion || h.slab_bytes != slab_bytes) { set_error("brg_index_attach: the buffer does not hold an index slab of this size"); return BRG_ERR_INVALID; }
    brg_index* ix = new brg_index();
    ix->ctx = ctx; ix->h = h; ix->slab = d_slab; ix->pooled = true; ix->view = make_view(h, d_slab);
    *out = ix;
    return BRG_OK;
}

// Replication inside ONE process (the Rust ExecutionPlan maps partitions to GPUs, interval_join.rs:442-550; the build side is
// shared by all partitions, joins/utils.rs:73-107): out[i] = a replica of `src` on the device of dst[i].  The slab travels
// device to device (NVLink peer copies where the topology has them) in a doubling tree: every GPU that already holds the
// slab feeds one that does not, so N replicas take ceil(log2(N + 1)) copy rounds instead of N copies out of one GPU.
extern "C" BRG_API int brg_index_replicate(const brg_index* src, brg_ctx* const* dst, int n_dst, brg_index** out) {
    if (!src || !src->ctx || (n_dst > 0 && (!dst || !out))) { set_error("brg_index_replicate: NULL argument"); return BRG_ERR_INVALID; }
    for (int i = 0; i < n_dst; i++) out[i] = nullptr;
    for (int i = 0; i < n_dst; i++)
        if (!dst[i]) { set_error("brg_index_replicate: NULL destination context"); return BRG_ERR_INVALID; }
    const uint64_t bytes = src->h.slab_bytes;
    std::vector<void*> slab((size_t)n_dst, nullptr);
    std::vector<cudaEvent_t> ready((size_t)n_dst + 1, nullptr);  // [0] = source, [i + 1] = replica i holds the bytes
    auto cleanup = [&](bool failed) {
        for (auto e : ready) if (e) cudaEventDestroy(e);
        if (failed) for (int i = 0; i < n_dst; i++) { if (slab[i]) { cudaSetDevice(dst[i]->device); cudaFree(slab[i]); } if (out[i]) { delete out[i]; out[i] = nullptr; } }
    };
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n_dst && e == cudaSuccess; i++) {
        e = cudaSetDevice(dst[i]->device);
        if (e == cudaSuccess) e = cudaMalloc(&slab[i], bytes);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ready[i + 1], cudaEventDisableTiming);
        if (e == cudaSuccess && dst[i]->device != src->ctx->device) {  // direct peer access where available (an error here only means "not needed / not possible")
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dst[i]->device, src->ctx->device) == cudaSuccess && can) cudaDeviceEnablePeerAccess(src->ctx->device, 0);
            cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaSetDevice(src->ctx->device);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ready[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ready[0], src->ctx->build_stream);
    // holder h (0 = the source, i + 1 = replica i) feeds replica t in the round in which the number of holders doubles
    for (int have = 1; have < n_dst + 1 && e == cudaSuccess; have *= 2) {
        for (int hsrc = 0; hsrc < have && hsrc + have < n_dst + 1 && e == cudaSuccess; hsrc++) {
            const int t = hsrc + have - 1;  // replica index
            const brg_ctx* from = hsrc == 0 ? src->ctx : dst[hsrc - 1];
            const void* from_p = hsrc == 0 ? src->slab : slab[hsrc - 1];
            e = cudaSetDevice(dst[t]->device);
            cudaStream_t st = dst[t]->build_stream;
            if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ready[hsrc], 0);
            if (e == cudaSuccess) e = cudaMemcpyPeerAsync(slab[t], dst[t]->device, from_p, from->device, bytes, st);
            if (e == cudaSuccess) e = cudaEventRecord(ready[t + 1], st);
        }
    }
    for (int i = 0; i < n_dst && e == cudaSuccess; i++) {
        e = cudaSetDevice(dst[i]->device);
        if (e == cudaSuccess) e = cudaStreamSynchronize(dst[i]->build_stream);
    }
    if (e != cudaSuccess) { set_error(std::string("brg_index_replicate: ") + cudaGetErrorString(e)); cudaGetLastError(); cleanup(true); return BRG_ERR_CUDA; }
    for (int i = 0; i < n_dst; i++) {
        brg_index* ix = new brg_index();
        ix->ctx = dst[i]; ix->h = src->h; ix->slab = slab[i]; ix->view = make_view(src->h, slab[i]);
        out[i] = ix;
    }
    cleanup(false);
    cudaSetDevice(src->ctx->device);
    return BRG_OK;
}

// ---------------------------------------------------------------------------------------------------
// device-buffer probes
// ---------------------------------------------------------------------------------------------------
static int check_probe(const brg_ctx* ctx, const brg_index* ix, uint32_t need, const char* what) {
    if (!ctx || !ix) { set_error(std::string(what) + ": NULL ctx/index"); return BRG_ERR_INVALID; }
    if ((ix->h.flags & need) == 0) { set_error(std::string(what) + ": index was built without the required structures"); return BRG_ERR_INVALID; }
    return 0;
}

extern "C" BRG_API int brg_count_overlaps_dev(brg_ctx* ctx, const brg_index* ix, const uint32_t* k, const int32_t* s,
                                      const int32_t* e, uint64_t n, int strict, int64_t* out, void* stream) {
    int rc = check_probe(ctx, ix, BRG_WANT_COUNT | BRG_WANT_NEAREST, "brg_count_overlaps");
    if (rc) return rc;
    return launch_count(ix->view, k, s, e, n, strict, out, (cudaStream_t)stream);
}
extern "C" BRG_API int brg_coverage_dev(brg_ctx* ctx, const brg_index* ix, const uint32_t* k, const int32_t* s, const int32_t* e,
                                uint64_t n, int strict, int64_t* out, void* stream) {
    int rc = check_probe(ctx, ix, BRG_WANT_COVERAGE, "brg_coverage");
    if (rc) return rc;
    return launch_coverage(ix->view, k, s, e, n, strict, out, (cudaStream_t)stream);
}
extern "C" BRG_API int brg_nearest_k1_dev(brg_ctx* ctx, const brg_index* ix, const uint32_t* k, const int32_t* s,
                                  const int32_t* e, uint64_t n, int include_overlaps, int strict, uint32_t* left_idx,
                                  uint32_t* valid_bits, int64_t* dist, void* stream) {
    int rc = check_probe(ctx, ix, BRG_WANT_NEAREST, "brg_nearest_k1");
    if (rc) return rc;
    if (n > 0xFFFFFFFFull) { set_error("right row index " + std::to_string(n - 1) + " exceeds UInt32 index capacity"); return BRG_ERR_OVERFLOW; }
    return launch_nearest_k1(ix->view, k, s, e, n, include_overlaps, strict, left_idx, valid_bits, dist, (cudaStream_t)stream);
}

static int pairs_scratch(Lane* l, uint64_t n) {
    int rc;
    if ((rc = l->scratch_a.ensure(std::max<uint64_t>(n, 1) * 4))) return rc;          // counts (u32)
    if ((rc = l->scratch_c.ensure((scan_scratch_words(std::max<uint64_t>(n, 1)) + 16) * 4))) return rc;
    return 0;
}

extern "C" BRG_API int brg_pairs_count_dev(brg_ctx* ctx, const brg_index* ix, const uint32_t* k, const int32_t* s,
                                   const int32_t* e, uint64_t n, uint64_t* d_offsets, void* stream) {
    int rc = check_probe(ctx, ix, BRG_WANT_PAIRS, "brg_pairs_count");
    if (rc) return rc;
    if (n > 0xFFFFFFFFull) { set_error("probe row index exceeds UInt32 index capacity"); return BRG_ERR_OVERFLOW; }
    // temporaries are stream-ordered allocations on the caller's stream: nothing to synchronise, as for every _dev entry point
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* counts = nullptr;
    uint64_t* scratch = nullptr;
    BRG_CUDA(cudaSetDevice(ctx->device));
    BRG_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&counts), std::max<uint64_t>(n, 1) * 4, st));
    cudaError_t e2 = cudaMallocAsync(reinterpret_cast<void**>(&scratch), (scan_scratch_words(std::max<uint64_t>(n, 1)) + 16) * 4, st);
    if (e2 != cudaSuccess) { cudaFreeAsync(counts, st); set_error(std::string("cudaMallocAsync: ") + cudaGetErrorString(e2)); return BRG_ERR_OOM; }
    rc = launch_pairs_count(ix->view, k, s, e, n, counts, st);
    if (!rc) rc = exclusive_scan_u32_to_u64(counts, d_offsets, n, scratch, st);
    cudaFreeAsync(counts, st);
    cudaFreeAsync(scratch, st);
    return rc;
}
extern "C" BRG_API int brg_pairs_emit_dev(brg_ctx* ctx, const brg_index* ix, const uint32_t* k, const int32_t* s,
                                  const int32_t* e, uint64_t n, const uint64_t* d_offsets, uint32_t* d_build_idx,
                                  uint32_t* d_probe_idx, void* stream) {
    int rc = check_probe(ctx, ix, BRG_WANT_PAIRS, "brg_pairs_emit");
    if (rc) return rc;
    return launch_pairs_emit(ix->view, k, s, e, n, d_offsets, d_build_idx, d_probe_idx, (cudaStream_t)stream);
}

extern "C" BRG_API uint32_t brg_pairs_chunk_rows(void) { return (uint32_t)TP_WTILE; }
extern "C" BRG_API int brg_pairs_chunk_offsets_dev(brg_ctx* ctx, const brg_index* ix, const uint32_t* k, const int32_t* s,
                                           const int32_t* e, uint64_t n, uint64_t* d_chunk_base, uint64_t* d_row_ranges,
                                           void* stream) {
    int rc = check_probe(ctx, ix, BRG_WANT_PAIRS, "brg_pairs_chunk_offsets");
    if (rc) return rc;
    if (n > 0xFFFFFFFFull) { set_error("probe row index exceeds UInt32 index capacity"); return BRG_ERR_OVERFLOW; }
    if (!d_chunk_base || (n && !d_row_ranges)) { set_error("brg_pairs_chunk_offsets: NULL output"); return BRG_ERR_INVALID; }
    return launch_pairs_chunk_offsets(ix->view, k, s, e, n, d_chunk_base, d_row_ranges, (cudaStream_t)stream);
}
extern "C" BRG_API int brg_pairs_fill_dev(brg_ctx* ctx, const brg_index* ix, const int32_t* s, uint64_t n,
                                  const uint64_t* d_chunk_base, const uint64_t* d_row_ranges, uint32_t* d_build_idx,
                                  uint32_t* d_probe_idx, void* stream) {
    int rc = check_probe(ctx, ix, BRG_WANT_PAIRS, "brg_pairs_fill");
    if (rc) return rc;
    return launch_pairs_fill(ix->view, s, n, d_chunk_base, d_row_ranges, d_build_idx, d_probe_idx, (cudaStream_t)stream);
}

extern "C" BRG_API uint32_t brg_pairs_join_chunk_rows(void) { return pairs_join_chunk_rows(); }
extern "C" BRG_API int brg_pairs_join_dev(brg_ctx* ctx, const brg_index* ix, const uint32_t* k, const int32_t* s, const int32_t* e,
                                          uint64_t n, uint64_t capacity, int flags, uint32_t* d_build_idx, uint32_t* d_probe_idx,
                                          uint64_t* d_total, uint64_t* d_chunk_slots, void* stream) {
    int rc = check_probe(ctx, ix, BRG_WANT_PAIRS, "brg_pairs_join");
    if (rc) return rc;
    if (n > 0xFFFFFFFFull) { set_error("probe row index exceeds UInt32 index capacity"); return BRG_ERR_OVERFLOW; }
    if (!d_total || (capacity && (!d_build_idx || !d_probe_idx))) { set_error("brg_pairs_join: NULL output"); return BRG_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    BRG_CUDA(cudaSetDevice(ctx->device));
    void* state = nullptr;  // the ticket, the cursor, the look-back words: a stream-ordered allocation, nothing to synchronise
    BRG_CUDA(cudaMallocAsync(&state, pairs_join_state_bytes(n), st));
    rc = launch_pairs_join(ix->view, k, s, e, n, capacity, (flags & BRG_JOIN_ORDERED) != 0, d_build_idx, d_probe_idx, d_total, d_chunk_slots,
                           state, st);
    cudaFreeAsync(state, st);
    return rc;
}

extern "C" BRG_API int brg_take_dev(brg_ctx* ctx, const void* d_src, uint32_t elem_bytes, const uint32_t* d_idx,
                            const uint32_t* d_valid_bits, uint64_t n, void* d_out, void* stream) {
    if (!ctx || (n && (!d_src || !d_idx || !d_out))) { set_error("brg_take: NULL argument"); return BRG_ERR_INVALID; }
    return launch_take(d_src, elem_bytes, d_idx, d_valid_bits, n, d_out, (cudaStream_t)stream) ? BRG_ERR_INVALID : BRG_OK;
}

// ---------------------------------------------------------------------------------------------------
// host-buffer probes with fixed-size outputs: chunked, double-buffered H2D / kernel / D2H pipeline
// ---------------------------------------------------------------------------------------------------
enum FixedOp { OP_COUNT, OP_COVERAGE, OP_NEAREST1 };

struct KeyRuns {  // run-length form of the probe key column (NULL run_key = not used)
    const uint32_t* run_key = nullptr;
    const uint64_t* run_end = nullptr;
    uint64_t n_runs = 0;
};

struct PosTypes {  // Arrow ingest: native width of the start / end value buffers, and where the first bad rows are reported
    int s_type = BRG_POS_INT32, e_type = BRG_POS_INT32;
    uint64_t* bad_start_row = nullptr;
    uint64_t* bad_end_row = nullptr;
};

// ---- pageable caller buffers ---------------------------------------------------------------------------------------
// cudaMemcpyAsync on pageable memory is staged by the driver on the calling thread: nothing overlaps and the copy runs at
// a fraction of the PCIe rate.  The Arrow buffers DataFusion hands over are ordinary heap memory, so run_fixed detects them
// and moves them through the lane's PINNED staging ring itself: the context's host threads copy chunk j + 1 in (and chunk
// j - 1 out) while chunk j is on the wire / in the kernel.
static bool is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
static void pool_memcpy(HostPool* pool, void* dst, const void* src, uint64_t bytes) {
    const uint64_t slice = 4ull << 20;
    const int parts = (int)std::min<uint64_t>((bytes + slice - 1) / slice, (uint64_t)pool->width() * 4);
    if (parts <= 1) { memcpy(dst, src, bytes); return; }
    const uint64_t per = ((bytes + parts - 1) / parts + 63) & ~63ull;
    pool->parallel_for(parts, [&](int i) {
        const uint64_t b = (uint64_t)i * per;
        if (b < bytes) memcpy((char*)dst + b, (const char*)src + b, std::min(per, bytes - b));
    });
}
// Runs of equal ids in key[0, m): at most `cap` of them are written as (id, end row relative to the chunk start + row0);
// returns the number of runs, or 0 when there are more than cap (the caller then sends the column itself).  This is the
// contig-cache loop of the reference's row loops (interval_tree.rs:258-265) done once per staging chunk: on a coordinate-
// sorted table a chunk holds a handful of runs, and 8 + 4 bytes per run cross PCIe instead of 4 bytes per row.
static uint32_t detect_key_runs(HostPool* pool, const uint32_t* key, uint64_t m, uint64_t row0, uint32_t cap, uint32_t* run_key,
                                uint64_t* run_end) {
    if (m == 0) return 0;
    const int parts = (int)std::min<uint64_t>((m + (1u << 20) - 1) >> 20, (uint64_t)pool->width() * 4);
    const uint64_t per = (m + parts - 1) / parts;
    std::vector<std::vector<uint64_t>> cuts((size_t)parts);  // rows i (relative) with key[i] != key[i - 1]
    std::atomic<uint32_t> total{0};
    pool->parallel_for(parts, [&](int t) {
        const uint64_t b = std::max<uint64_t>((uint64_t)t * per, 1), e = std::min(m, ((uint64_t)t + 1) * per);
        std::vector<uint64_t>& c = cuts[(size_t)t];
        for (uint64_t blk = b; blk < e && total.load(std::memory_order_relaxed) <= cap; blk += 4096) {
            const uint64_t be = std::min(e, blk + 4096);
            uint32_t acc = 0;
            for (uint64_t i = blk; i < be; i++) acc |= key[i] ^ key[i - 1];  // (vectorises; almost every block has no boundary)
            if (!acc) continue;
            for (uint64_t i = blk; i < be; i++)
                if (key[i] != key[i - 1]) { c.push_back(i); total.fetch_add(1, std::memory_order_relaxed); }
        }
    });
    if (total.load() + 1 > cap) return 0;
    uint32_t r = 0;
    uint64_t begin = 0;
    for (auto& c : cuts)
        for (uint64_t cut : c) { run_key[r] = key[begin]; run_end[r] = row0 + cut; r++; begin = cut; }
    run_key[r] = key[begin]; run_end[r] = row0 + m; r++;
    return r;
}
constexpr uint32_t kMaxChunkRuns = 1024;

static int run_fixed(brg_ctx* ctx, const brg_index* ix, FixedOp op, const uint32_t* key, const void* start_v,
                     const void* end_v, uint64_t n, int include_overlaps, int strict, int64_t* out_i64,
                     uint32_t* out_left, uint8_t* out_bits, int64_t* out_dist, KeyRuns runs = KeyRuns(), PosTypes pt = PosTypes()) {
    if (pt.bad_start_row) *pt.bad_start_row = ~0ull;
    if (pt.bad_end_row) *pt.bad_end_row = ~0ull;
    if (n == 0) return BRG_OK;
    if (!start_v || !end_v) { set_error("probe: NULL column"); return BRG_ERR_INVALID; }
    if (pt.s_type < BRG_POS_INT32 || pt.s_type > BRG_POS_UINT64 || pt.e_type < BRG_POS_INT32 || pt.e_type > BRG_POS_UINT64) {
        set_error("probe: unknown position column type");
        return BRG_ERR_INVALID;
    }
    const char* start = static_cast<const char*>(start_v);
    const char* end = static_cast<const char*>(end_v);
    const uint64_t ws = pos_width(pt.s_type), we = pos_width(pt.e_type);
    const bool narrow_s = pt.s_type != BRG_POS_INT32, narrow_e = pt.e_type != BRG_POS_INT32;
    const bool by_runs = runs.run_key != nullptr;
    if (by_runs) {
        if (!runs.run_end || runs.n_runs == 0 || runs.n_runs > 0xFFFFFFFFull || runs.run_end[runs.n_runs - 1] != n) {
            set_error("probe: key runs must be non-empty and end at the row count");
            return BRG_ERR_INVALID;
        }
    }
    BRG_CUDA(cudaSetDevice(ctx->device));
    LaneGuard g(ctx);
    int rc;
    if ((rc = g.acquire())) return rc;
    Lane* l = g.l;
    // which caller buffers are pageable (BRG_STAGE=0 leaves them to the driver: the measured alternative)
    static const bool stage_on = !(getenv("BRG_STAGE") && getenv("BRG_STAGE")[0] == '0');
    static const bool runs_on = !(getenv("BRG_CHUNK_RUNS") && getenv("BRG_CHUNK_RUNS")[0] == '0');
    const bool pg_s = stage_on && is_pageable(start), pg_e = stage_on && is_pageable(end);
    const bool pg_o0 = stage_on && is_pageable(op == OP_NEAREST1 ? (const void*)out_left : (const void*)out_i64);
    const bool pg_o1 = stage_on && op == OP_NEAREST1 && is_pageable(out_bits);
    const bool pg_o2 = stage_on && op == OP_NEAREST1 && out_dist && is_pageable(out_dist);
    const bool pg_key = stage_on && key && !by_runs && is_pageable(key);
    // BRG_NARROW_OUT=1 (experiment): Int64 count / coverage values fit 32 bits (< 2^31 build rows; the coverage accumulator is
    // i32): send 4 bytes per row over PCIe and let the host threads widen them into the caller's Int64 buffer
    static const bool narrow_env = getenv("BRG_NARROW_OUT") && getenv("BRG_NARROW_OUT")[0] == '1';
    const bool narrow_out = narrow_env && op != OP_NEAREST1 && ix->h.n < 0x80000000u;
    const bool staged_out = pg_o0 || pg_o1 || pg_o2 || narrow_out;
    const bool chunk_runs = runs_on && key && !by_runs;  // look for runs in every chunk of a per-row key column
    HostPool* pool = (pg_s || pg_e || pg_key || staged_out || chunk_runs) ? host_pool(ctx) : nullptr;
    auto widen = [&](int64_t* dst, const int32_t* src, uint64_t m) {
        const int parts = (int)std::min<uint64_t>((m + (1u << 20) - 1) >> 20, (uint64_t)pool->width() * 4);
        const uint64_t per = (m + std::max(parts, 1) - 1) / std::max(parts, 1);
        pool->parallel_for(std::max(parts, 1), [&](int t) {
            const uint64_t b = (uint64_t)t * per, e = std::min(m, b + per);
            for (uint64_t i = b; i < e; i++) dst[i] = (int64_t)src[i];
        });
    };

    // everything below is asynchronous on the lane's three streams: a failure must drain them before the lane (and the
    // caller's buffers) are released
    auto body = [&]() -> int {
        int rc;
        if (by_runs) {  // the runs travel once (a few bytes per contig); the per-row key column is expanded on the device
            if ((rc = l->scratch_a.ensure(runs.n_runs * 4)) || (rc = l->scratch_b.ensure(runs.n_runs * 8))) return rc;
            BRG_CUDA(cudaMemcpyAsync(l->scratch_a.p, runs.run_key, runs.n_runs * 4, cudaMemcpyHostToDevice, l->k));
            BRG_CUDA(cudaMemcpyAsync(l->scratch_b.p, runs.run_end, runs.n_runs * 8, cudaMemcpyHostToDevice, l->k));
        }
        const bool dev_key = key != nullptr || by_runs;
        // chunk rows are a multiple of 32 so that validity bitmap words never straddle chunks
        const uint64_t C = std::min<uint64_t>((ctx->chunk_rows + 31) & ~31ull, (n + 31) & ~31ull);
        if (narrow_s || narrow_e) {
            if ((rc = l->bad_rows.ensure(16))) return rc;
            BRG_CUDA(cudaMemsetAsync(l->bad_rows.p, 0xFF, 16, l->k));
        }
        for (int b = 0; b < Lane::NBUF; b++) {
            if (dev_key && (rc = l->in_key[b].ensure(C * 4))) return rc;
            if ((rc = l->in_s[b].ensure(C * 4)) || (rc = l->in_e[b].ensure(C * 4))) return rc;
            if (narrow_s && (rc = l->raw_s[b].ensure(C * ws))) return rc;
            if (narrow_e && (rc = l->raw_e[b].ensure(C * we))) return rc;
            if (op == OP_NEAREST1) {
                if ((rc = l->out0[b].ensure(C * 4)) || (rc = l->out1[b].ensure(C / 8 + 4))) return rc;
                if (out_dist && (rc = l->out2[b].ensure(C * 8))) return rc;
            } else {
                if ((rc = l->out0[b].ensure(C * 8))) return rc;
            }
            if (pg_s && (rc = l->stage_s[b].ensure(C * ws))) return rc;
            if (pg_e && (rc = l->stage_e[b].ensure(C * we))) return rc;
            if (pg_key && (rc = l->stage_key[b].ensure(C * 4))) return rc;
            if ((pg_o0 || narrow_out) && (rc = l->stage_o0[b].ensure(C * 8))) return rc;
            if (narrow_out && (rc = l->out1[b].ensure(C * 4))) return rc;
            if (pg_o1 && (rc = l->stage_o1[b].ensure(C / 8 + 4))) return rc;
            if (pg_o2 && (rc = l->stage_o2[b].ensure(C * 8))) return rc;
            if (chunk_runs && ((rc = l->stage_runs[b].ensure((uint64_t)kMaxChunkRuns * 12)) || (rc = l->runs_key[b].ensure((uint64_t)kMaxChunkRuns * 4)) ||
                               (rc = l->runs_end[b].ensure((uint64_t)kMaxChunkRuns * 8))))
                return rc;
        }
        // chunk schedule: full chunks of C rows.  (Tapering the ends -- C/8, C/4, C/2 ... -- to shorten the first H2D and the
        // last D2H copy, which nothing overlaps, was measured and is slower: 24.2 ms vs 22.4 ms per 125M probes; so are smaller
        // chunks throughout, profiles/r1_e2e_chunk_sweep.txt: the per-chunk cost outweighs the shorter pipeline fill.)
        std::vector<uint64_t> sizes, offs;
        for (uint64_t left = n, at = 0; left;) { const uint64_t m = std::min(C, left); sizes.push_back(m); offs.push_back(at); at += m; left -= m; }
        // drains chunk j's staged results into the caller's (pageable) output buffers
        auto drain = [&](uint64_t j) -> int {
            const int b = (int)(j % Lane::NBUF);
            const uint64_t m = sizes[j], done = offs[j];
            BRG_CUDA(cudaEventSynchronize(l->ev_out[b]));
            if (op == OP_NEAREST1) {
                if (pg_o0) pool_memcpy(pool, out_left + done, l->stage_o0[b].p, m * 4);
                if (pg_o1) pool_memcpy(pool, out_bits + done / 8, l->stage_o1[b].p, ((m + 31) / 32) * 4);
                if (pg_o2) pool_memcpy(pool, out_dist + done, l->stage_o2[b].p, m * 8);
            } else if (narrow_out) {
                widen(out_i64 + done, l->stage_o0[b].as<int32_t>(), m);
            } else if (pg_o0) {
                pool_memcpy(pool, out_i64 + done, l->stage_o0[b].p, m * 8);
            }
            return 0;
        };
        for (uint64_t j = 0; j < sizes.size(); j++) {
            const int b = (int)(j % Lane::NBUF);
            const uint64_t m = sizes[j], done = offs[j];
            // ---- host side of the input: staging copies and the key-run scan (they overlap the chunks already enqueued) ----
            if ((pg_s || pg_e || pg_key || chunk_runs) && j >= (uint64_t)Lane::NBUF)
                BRG_CUDA(cudaEventSynchronize(l->ev_in[b]));  // the H2D copies that read this staging buffer are done
            uint32_t n_chunk_runs = 0;
            if (chunk_runs) {
                uint32_t* rk = l->stage_runs[b].as<uint32_t>();
                uint64_t* re = reinterpret_cast<uint64_t*>(l->stage_runs[b].as<char>() + (uint64_t)kMaxChunkRuns * 4);
                n_chunk_runs = detect_key_runs(pool, key + done, m, 0, kMaxChunkRuns, rk, re);
            }
            const bool send_key = key && !by_runs && n_chunk_runs == 0;
            const void* src_s = start + done * ws;
            const void* src_e = end + done * we;
            const void* src_k = key ? (const void*)(key + done) : nullptr;
            if (pg_s) { pool_memcpy(pool, l->stage_s[b].p, src_s, m * ws); src_s = l->stage_s[b].p; }
            if (pg_e) { pool_memcpy(pool, l->stage_e[b].p, src_e, m * we); src_e = l->stage_e[b].p; }
            if (pg_key && send_key) { pool_memcpy(pool, l->stage_key[b].p, src_k, m * 4); src_k = l->stage_key[b].p; }
            // ---- device side -----------------------------------------------------------------------------------------------
            // inputs of buffer b may be overwritten once the kernel that read them has finished
            BRG_CUDA(cudaStreamWaitEvent(l->h2d, l->ev_k[b], 0));
            if (send_key) BRG_CUDA(cudaMemcpyAsync(l->in_key[b].p, src_k, m * 4, cudaMemcpyHostToDevice, l->h2d));
            if (n_chunk_runs) {
                BRG_CUDA(cudaMemcpyAsync(l->runs_key[b].p, l->stage_runs[b].p, (uint64_t)n_chunk_runs * 4, cudaMemcpyHostToDevice, l->h2d));
                BRG_CUDA(cudaMemcpyAsync(l->runs_end[b].p, l->stage_runs[b].as<char>() + (uint64_t)kMaxChunkRuns * 4, (uint64_t)n_chunk_runs * 8,
                                         cudaMemcpyHostToDevice, l->h2d));
            }
            BRG_CUDA(cudaMemcpyAsync(narrow_s ? l->raw_s[b].p : l->in_s[b].p, src_s, m * ws, cudaMemcpyHostToDevice, l->h2d));
            BRG_CUDA(cudaMemcpyAsync(narrow_e ? l->raw_e[b].p : l->in_e[b].p, src_e, m * we, cudaMemcpyHostToDevice, l->h2d));
            BRG_CUDA(cudaEventRecord(l->ev_in[b], l->h2d));
            // kernel: after its inputs landed and after the previous results of buffer b left the device
            BRG_CUDA(cudaStreamWaitEvent(l->k, l->ev_in[b], 0));
            BRG_CUDA(cudaStreamWaitEvent(l->k, l->ev_out[b], 0));
            const uint32_t* dk = dev_key ? l->in_key[b].as<uint32_t>() : nullptr;
            if (by_runs) {  // (the k stream already waited for the previous kernel that read this buffer)
                expand_runs_kernel<<<(unsigned)((m + 255) / 256), 256, 0, l->k>>>(l->scratch_a.as<uint32_t>(),
                                                                                  l->scratch_b.as<unsigned long long>(),
                                                                                  (uint32_t)runs.n_runs, done, m, l->in_key[b].as<uint32_t>());
                g_launches.fetch_add(1);
            } else if (n_chunk_runs) {
                expand_runs_kernel<<<(unsigned)((m + 255) / 256), 256, 0, l->k>>>(l->runs_key[b].as<uint32_t>(),
                                                                                  l->runs_end[b].as<unsigned long long>(), n_chunk_runs, 0, m,
                                                                                  l->in_key[b].as<uint32_t>());
                g_launches.fetch_add(1);
            }
            if (narrow_s && (rc = launch_narrow(l->raw_s[b].p, pt.s_type, m, done, l->in_s[b].as<int32_t>(), l->bad_rows.as<unsigned long long>(), l->k))) return rc;
            if (narrow_e && (rc = launch_narrow(l->raw_e[b].p, pt.e_type, m, done, l->in_e[b].as<int32_t>(), l->bad_rows.as<unsigned long long>() + 1, l->k))) return rc;
            if (op == OP_COUNT)
                rc = launch_count(ix->view, dk, l->in_s[b].as<int32_t>(), l->in_e[b].as<int32_t>(), m, strict, l->out0[b].as<int64_t>(), l->k);
            else if (op == OP_COVERAGE)
                rc = launch_coverage(ix->view, dk, l->in_s[b].as<int32_t>(), l->in_e[b].as<int32_t>(), m, strict, l->out0[b].as<int64_t>(), l->k);
            else
                rc = launch_nearest_k1(ix->view, dk, l->in_s[b].as<int32_t>(), l->in_e[b].as<int32_t>(), m, include_overlaps, strict,
                                       l->out0[b].as<uint32_t>(), l->out1[b].as<uint32_t>(), out_dist ? l->out2[b].as<int64_t>() : nullptr, l->k);
            if (rc) return rc;
            if (narrow_out) {
                narrow_i64_kernel<<<(unsigned)((m + 255) / 256), 256, 0, l->k>>>(l->out0[b].as<long long>(), m, l->out1[b].as<int32_t>());
                g_launches.fetch_add(1);
            }
            BRG_CUDA(cudaEventRecord(l->ev_k[b], l->k));
            BRG_CUDA(cudaStreamWaitEvent(l->d2h, l->ev_k[b], 0));
            if (op == OP_NEAREST1) {
                BRG_CUDA(cudaMemcpyAsync(pg_o0 ? l->stage_o0[b].p : (void*)(out_left + done), l->out0[b].p, m * 4, cudaMemcpyDeviceToHost, l->d2h));
                BRG_CUDA(cudaMemcpyAsync(pg_o1 ? l->stage_o1[b].p : (void*)(out_bits + done / 8), l->out1[b].p, ((m + 31) / 32) * 4, cudaMemcpyDeviceToHost, l->d2h));
                if (out_dist) BRG_CUDA(cudaMemcpyAsync(pg_o2 ? l->stage_o2[b].p : (void*)(out_dist + done), l->out2[b].p, m * 8, cudaMemcpyDeviceToHost, l->d2h));
            } else if (narrow_out) {
                BRG_CUDA(cudaMemcpyAsync(l->stage_o0[b].p, l->out1[b].p, m * 4, cudaMemcpyDeviceToHost, l->d2h));
            } else {
                BRG_CUDA(cudaMemcpyAsync(pg_o0 ? l->stage_o0[b].p : (void*)(out_i64 + done), l->out0[b].p, m * 8, cudaMemcpyDeviceToHost, l->d2h));
            }
            BRG_CUDA(cudaEventRecord(l->ev_out[b], l->d2h));
            // ---- host side of the output: the previous chunk's results leave the staging ring while this one is in flight ----
            if (staged_out && j >= 1 && (rc = drain(j - 1))) return rc;
        }
        if (staged_out && (rc = drain(sizes.size() - 1))) return rc;
        uint64_t bad[2] = {~0ull, ~0ull};
        if (narrow_s || narrow_e) BRG_CUDA(cudaMemcpyAsync(bad, l->bad_rows.p, 16, cudaMemcpyDeviceToHost, l->k));
        BRG_CUDA(cudaStreamSynchronize(l->d2h));
        BRG_CUDA(cudaStreamSynchronize(l->k));
        BRG_CUDA(cudaStreamSynchronize(l->h2d));
        if (bad[0] != ~0ull || bad[1] != ~0ull) {  // array_utils.rs:94-98 (the start column is resolved first)
            if (pt.bad_start_row) *pt.bad_start_row = bad[0];
            if (pt.bad_end_row) *pt.bad_end_row = bad[1];
            const bool in_start = bad[0] != ~0ull;
            const uint64_t row = in_start ? bad[0] : bad[1];
            const int t = in_start ? pt.s_type : pt.e_type;
            const char* col = in_start ? start : end;
            std::string v;
            if (t == BRG_POS_INT64) v = std::to_string(reinterpret_cast<const int64_t*>(col)[row]);
            else if (t == BRG_POS_UINT32) v = std::to_string(reinterpret_cast<const uint32_t*>(col)[row]);
            else v = std::to_string(reinterpret_cast<const uint64_t*>(col)[row]);
            set_error("coordinate value " + v + " at row " + std::to_string(row) + " overflows i32 (max 2147483647)");
            return BRG_ERR_OVERFLOW;
        }
        return BRG_OK;
    };
    rc = body();
    if (rc) {  // in-flight copies / kernels still use the lane's and the caller's buffers: drain before anyone reuses them
        const std::string msg = brg_last_error();
        cudaStreamSynchronize(l->h2d); cudaStreamSynchronize(l->k); cudaStreamSynchronize(l->d2h);
        cudaGetLastError();
        set_error(msg);
    }
    return rc;
}

extern "C" BRG_API int brg_count_overlaps(brg_ctx* ctx, const brg_index* ix, const uint32_t* key, const int32_t* start,
                                  const int32_t* end, uint64_t n, int strict, int64_t* out) {
    int rc = check_probe(ctx, ix, BRG_WANT_COUNT | BRG_WANT_NEAREST, "brg_count_overlaps");
    if (rc) return rc;
    if (n && !out) { set_error("brg_count_overlaps: NULL output"); return BRG_ERR_INVALID; }
    return run_fixed(ctx, ix, OP_COUNT, key, start, end, n, 0, strict, out, nullptr, nullptr, nullptr);
}
extern "C" BRG_API int brg_count_overlaps_runs(brg_ctx* ctx, const brg_index* ix, const uint32_t* run_key, const uint64_t* run_end,
                                       uint64_t n_runs, const int32_t* start, const int32_t* end, uint64_t n, int strict,
                                       int64_t* out) {
    int rc = check_probe(ctx, ix, BRG_WANT_COUNT | BRG_WANT_NEAREST, "brg_count_overlaps_runs");
    if (rc) return rc;
    if (n && (!out || !run_key)) { set_error("brg_count_overlaps_runs: NULL argument"); return BRG_ERR_INVALID; }
    KeyRuns kr; kr.run_key = run_key; kr.run_end = run_end; kr.n_runs = n_runs;
    return run_fixed(ctx, ix, OP_COUNT, nullptr, start, end, n, 0, strict, out, nullptr, nullptr, nullptr, kr);
}
extern "C" BRG_API int brg_coverage_runs(brg_ctx* ctx, const brg_index* ix, const uint32_t* run_key, const uint64_t* run_end,
                                 uint64_t n_runs, const int32_t* start, const int32_t* end, uint64_t n, int strict, int64_t* out) {
    int rc = check_probe(ctx, ix, BRG_WANT_COVERAGE, "brg_coverage_runs");
    if (rc) return rc;
    if (n && (!out || !run_key)) { set_error("brg_coverage_runs: NULL argument"); return BRG_ERR_INVALID; }
    KeyRuns kr; kr.run_key = run_key; kr.run_end = run_end; kr.n_runs = n_runs;
    return run_fixed(ctx, ix, OP_COVERAGE, nullptr, start, end, n, 0, strict, out, nullptr, nullptr, nullptr, kr);
}
extern "C" BRG_API int brg_nearest_k1_runs(brg_ctx* ctx, const brg_index* ix, const uint32_t* run_key, const uint64_t* run_end,
                                   uint64_t n_runs, const int32_t* start, const int32_t* end, uint64_t n, int include_overlaps,
                                   int strict, uint32_t* out_left, uint8_t* out_bits, int64_t* out_dist) {
    int rc = check_probe(ctx, ix, BRG_WANT_NEAREST, "brg_nearest_k1_runs");
    if (rc) return rc;
    if (n > 0xFFFFFFFFull) { set_error("right row index " + std::to_string(n - 1) + " exceeds UInt32 index capacity"); return BRG_ERR_OVERFLOW; }
    if (n && (!out_left || !out_bits || !run_key)) { set_error("brg_nearest_k1_runs: NULL argument"); return BRG_ERR_INVALID; }
    KeyRuns kr; kr.run_key = run_key; kr.run_end = run_end; kr.n_runs = n_runs;
    return run_fixed(ctx, ix, OP_NEAREST1, nullptr, start, end, n, include_overlaps, strict, nullptr, out_left, out_bits, out_dist, kr);
}
// Arrow ingest entry points: the position value buffers arrive in their native Arrow width
static PosTypes pos_types(int s_type, int e_type, uint64_t* bad_rows) {
    PosTypes pt;
    pt.s_type = s_type; pt.e_type = e_type;
    if (bad_rows) { pt.bad_start_row = bad_rows; pt.bad_end_row = bad_rows + 1; }
    return pt;
}
extern "C" BRG_API int brg_count_overlaps_arrow(brg_ctx* ctx, const brg_index* ix, const uint32_t* run_key, const uint64_t* run_end,
                                        uint64_t n_runs, const void* start, int start_type, const void* end, int end_type,
                                        uint64_t n, int strict, int64_t* out, uint64_t* bad_rows) {
    int rc = check_probe(ctx, ix, BRG_WANT_COUNT | BRG_WANT_NEAREST, "brg_count_overlaps_arrow");
    if (rc) return rc;
    if (n && (!out || !run_key)) { set_error("brg_count_overlaps_arrow: NULL argument"); return BRG_ERR_INVALID; }
    KeyRuns kr; kr.run_key = run_key; kr.run_end = run_end; kr.n_runs = n_runs;
    return run_fixed(ctx, ix, OP_COUNT, nullptr, start, end, n, 0, strict, out, nullptr, nullptr, nullptr, kr, pos_types(start_type, end_type, bad_rows));
}
extern "C" BRG_API int brg_coverage_arrow(brg_ctx* ctx, const brg_index* ix, const uint32_t* run_key, const uint64_t* run_end,
                                  uint64_t n_runs, const void* start, int start_type, const void* end, int end_type, uint64_t n,
                                  int strict, int64_t* out, uint64_t* bad_rows) {
    int rc = check_probe(ctx, ix, BRG_WANT_COVERAGE, "brg_coverage_arrow");
    if (rc) return rc;
    if (n && (!out || !run_key)) { set_error("brg_coverage_arrow: NULL argument"); return BRG_ERR_INVALID; }
    KeyRuns kr; kr.run_key = run_key; kr.run_end = run_end; kr.n_runs = n_runs;
    return run_fixed(ctx, ix, OP_COVERAGE, nullptr, start, end, n, 0, strict, out, nullptr, nullptr, nullptr, kr, pos_types(start_type, end_type, bad_rows));
}
extern "C" BRG_API int brg_nearest_k1_arrow(brg_ctx* ctx, const brg_index* ix, const uint32_t* run_key, const uint64_t* run_end,
                                    uint64_t n_runs, const void* start, int start_type, const void* end, int end_type, uint64_t n,
                                    int include_overlaps, int strict, uint32_t* out_left, uint8_t* out_bits, int64_t* out_dist,
                                    uint64_t* bad_rows) {
    int rc = check_probe(ctx, ix, BRG_WANT_NEAREST, "brg_nearest_k1_arrow");
    if (rc) return rc;
    if (n > 0xFFFFFFFFull) { set_error("right row index " + std::to_string(n - 1) + " exceeds UInt32 index capacity"); return BRG_ERR_OVERFLOW; }
    if (n && (!out_left || !out_bits || !run_key)) { set_error("brg_nearest_k1_arrow: NULL argument"); return BRG_ERR_INVALID; }
    KeyRuns kr; kr.run_key = run_key; kr.run_end = run_end; kr.n_runs = n_runs;
    return run_fixed(ctx, ix, OP_NEAREST1, nullptr, start, end, n, include_overlaps, strict, nullptr, out_left, out_bits, out_dist, kr,
                     pos_types(start_type, end_type, bad_rows));
}
extern "C" BRG_API int brg_coverage(brg_ctx* ctx, const brg_index* ix, const uint32_t* key, const int32_t* start,
                            const int32_t* end, uint64_t n, int strict, int64_t* out) {
    int rc = check_probe(ctx, ix, BRG_WANT_COVERAGE, "brg_coverage");
    if (rc) return rc;
    if (n && !out) { set_error("brg_coverage: NULL output"); return BRG_ERR_INVALID; }
    return run_fixed(ctx, ix, OP_COVERAGE, key, start, end, n, 0, strict, out, nullptr, nullptr, nullptr);
}
extern "C" BRG_API int brg_nearest_k1(brg_ctx* ctx, const brg_index* ix, const uint32_t* key, const int32_t* start,
                              const int32_t* end, uint64_t n, int include_overlaps, int strict, uint32_t* out_left,
                              uint8_t* out_bits, int64_t* out_dist) {
    int rc = check_probe(ctx, ix, BRG_WANT_NEAREST, "brg_nearest_k1");
    if (rc) return rc;
    if (n > 0xFFFFFFFFull) { set_error("right row index " + std::to_string(n - 1) + " exceeds UInt32 index capacity"); return BRG_ERR_OVERFLOW; }
    if (n && (!out_left || !out_bits)) { set_error("brg_nearest_k1: NULL output"); return BRG_ERR_INVALID; }
    return run_fixed(ctx, ix, OP_NEAREST1, key, start, end, n, include_overlaps, strict, nullptr, out_left, out_bits, out_dist);
}

// ---------------------------------------------------------------------------------------------------
// variable-length results
// ---------------------------------------------------------------------------------------------------
struct brg_rows {
    brg_ctx* ctx = nullptr;
    uint64_t len = 0;
    PinnedBlock blk_build{}, blk_probe{}, blk_bits{}, blk_dist{}, blk_ps{}, blk_pe{};
};
extern "C" BRG_API uint64_t brg_rows_len(const brg_rows* r) { return r ? r->len : 0; }
extern "C" BRG_API const uint32_t* brg_rows_build_idx(const brg_rows* r) { return r ? (const uint32_t*)r->blk_build.p : nullptr; }
extern "C" BRG_API const uint32_t* brg_rows_probe_idx(const brg_rows* r) { return r ? (const uint32_t*)r->blk_probe.p : nullptr; }
extern "C" BRG_API const uint8_t* brg_rows_valid_bits(const brg_rows* r) { return r ? (const uint8_t*)r->blk_bits.p : nullptr; }
extern "C" BRG_API const int64_t* brg_rows_distance(const brg_rows* r) { return r ? (const int64_t*)r->blk_dist.p : nullptr; }
extern "C" BRG_API const int32_t* brg_rows_piece_start(const brg_rows* r) { return r ? (const int32_t*)r->blk_ps.p : nullptr; }
extern "C" BRG_API const int32_t* brg_rows_piece_end(const brg_rows* r) { return r ? (const int32_t*)r->blk_pe.p : nullptr; }
extern "C" BRG_API void brg_rows_release(brg_rows* r) {
    if (!r) return;
    pinned_put(r->ctx, r->blk_ps); pinned_put(r->ctx, r->blk_pe);
    pinned_put(r->ctx, r->blk_build); pinned_put(r->ctx, r->blk_probe);
    pinned_put(r->ctx, r->blk_bits); pinned_put(r->ctx, r->blk_dist);
    ctx_unref(r->ctx);
    delete r;
}

// copies the probe columns of rows [0,n) to the lane's input buffers (set 0) on the lane's kernel stream
static int upload_probe(Lane* l, const uint32_t* key, const int32_t* start, const int32_t* end, uint64_t n) {
    int rc;
    const uint64_t b = std::max<uint64_t>(n, 1) * 4;
    if (key && (rc = l->in_key[0].ensure(b))) return rc;
    if ((rc = l->in_s[0].ensure(b)) || (rc = l->in_e[0].ensure(b))) return rc;
    if (n) {
        if (key) BRG_CUDA(cudaMemcpyAsync(l->in_key[0].p, key, n * 4, cudaMemcpyHostToDevice, l->k));
        BRG_CUDA(cudaMemcpyAsync(l->in_s[0].p, start, n * 4, cudaMemcpyHostToDevice, l->k));
        BRG_CUDA(cudaMemcpyAsync(l->in_e[0].p, end, n * 4, cudaMemcpyHostToDevice, l->k));
    }
    return 0;
}

extern "C" BRG_API int brg_overlap_pairs(brg_ctx* ctx, const brg_index* ix, const uint32_t* key, const int32_t* start,
                                 const int32_t* end, uint64_t n, int mode, uint64_t max_rows, brg_rows** out,
                                 uint64_t* consumed) {
    int rc = check_probe(ctx, ix, BRG_WANT_PAIRS, "brg_overlap_pairs");
    if (rc) return rc;
    if (!out) { set_error("brg_overlap_pairs: NULL out"); return BRG_ERR_INVALID; }
    *out = nullptr;
    if (mode < BRG_PAIRS_INNER || mode > BRG_PAIRS_INNER_ANY_ORDER) { set_error("brg_overlap_pairs: bad mode"); return BRG_ERR_INVALID; }
    // The one-launch join is the fast one for coordinate-sorted probe tables; an unsorted one sends every chunk down its slow path
    // (2.9 ms against 1.26 ms for the chunked form on 10M shuffled rows).  Both are exact, so a sample decides: 1024 adjacent row
    // pairs spread over the table, mostly in (key, start) order or not.
    auto looks_sorted = [&]() {
        if (n < 4096) return true;
        const uint64_t stride = n / 1024;
        uint32_t bad = 0;
        for (uint64_t i = 0; i + 1 < n && i / stride < 1024; i += stride) {
            const uint32_t k0 = key ? key[i] : 0u, k1 = key ? key[i + 1] : 0u;
            bad += (k0 > k1 || (k0 == k1 && start[i] > start[i + 1])) ? 1u : 0u;
        }
        return bad < 64;
    };
    const bool want_any_order = mode == BRG_PAIRS_INNER_ANY_ORDER && max_rows == 0;
    if (mode == BRG_PAIRS_INNER_ANY_ORDER) mode = BRG_PAIRS_INNER;  // (with max_rows: the ordered pipeline, whose prefix can be cut)
    if (n > 0xFFFFFFFFull) { set_error("probe row index " + std::to_string(n - 1) + " exceeds UInt32 index capacity"); return BRG_ERR_OVERFLOW; }
    if (n && (!start || !end)) { set_error("brg_overlap_pairs: NULL column"); return BRG_ERR_INVALID; }
    const bool any_order = want_any_order && looks_sorted();
    BRG_CUDA(cudaSetDevice(ctx->device));
    brg_rows* r = new brg_rows();
    r->ctx = ctx;
    ctx->refs.fetch_add(1);
    if (consumed) *consumed = n;
    if (n == 0) { *out = r; return BRG_OK; }
    LaneGuard g(ctx);
    if ((rc = g.acquire())) { brg_rows_release(r); return rc; }
    Lane* l = g.l;
    cudaStream_t st = l->k;
    auto fail = [&](int code) { cudaStreamSynchronize(st); brg_rows_release(r); return code; };
    if ((rc = upload_probe(l, key, start, end, n))) return fail(rc);
    if ((rc = pairs_scratch(l, n))) return fail(rc);
    if ((rc = l->scratch_b.ensure((n + 1) * 8))) return fail(rc);  // offsets u64
    const uint32_t* dk = key ? l->in_key[0].as<uint32_t>() : nullptr;
    const int32_t* ds = l->in_s[0].as<int32_t>();
    const int32_t* de = l->in_e[0].as<int32_t>();
    if (any_order) {
        // the join in one launch (brg_join.cu) into the lane's output buffers; their size is the capacity.  A first call guesses
        // 6 pairs per probe row; a call that finds more grows the buffers and runs again (the total is exact either way).
        if ((rc = l->scratch_b.ensure(pairs_join_state_bytes(n) + 16))) return fail(rc);
        uint64_t* d_total = l->scratch_b.as<uint64_t>();
        void* state = l->scratch_b.as<uint64_t>() + 2;
        uint64_t cap = std::min<uint64_t>(l->out0[0].bytes, l->out1[0].bytes) / 4;
        if (cap < n * 6) {
            cap = std::max<uint64_t>(n * 6, 1024);
            if ((rc = l->out0[0].ensure(cap * 4)) || (rc = l->out1[0].ensure(cap * 4))) return fail(rc);
        }
        uint64_t total = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            if ((rc = launch_pairs_join(ix->view, dk, ds, de, n, cap, 0, l->out1[0].as<uint32_t>(), l->out0[0].as<uint32_t>(), d_total, nullptr,
                                        state, st)))
                return fail(rc);
            cudaError_t e = cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { set_error(std::string("pairs total: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
            if (total <= cap) break;
            cap = total;
            if ((rc = l->out0[0].ensure(cap * 4)) || (rc = l->out1[0].ensure(cap * 4))) return fail(rc);
        }
        r->len = total;
        if ((rc = pinned_get(ctx, total * 4, &r->blk_probe)) || (rc = pinned_get(ctx, total * 4, &r->blk_build))) return fail(rc);
        if (total) {
            cudaError_t e = cudaMemcpyAsync(r->blk_probe.p, l->out0[0].p, total * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(r->blk_build.p, l->out1[0].p, total * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { set_error(std::string("pairs D2H: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
        }
        *out = r;
        return BRG_OK;
    }
    if (mode == BRG_PAIRS_INNER && max_rows == 0) {
        // chunked count -> scan -> emit (brg_pairs.cu): offsets per 256-row chunk, staged coalesced emission
        const uint64_t n_chunks = pairs_chunk_count(n);
        if ((rc = l->scratch_b.ensure((n_chunks + 1) * 8)) || (rc = l->scratch_a.ensure(n * 8))) return fail(rc);
        uint64_t* chunk_base = l->scratch_b.as<uint64_t>();
        uint64_t* row_rng = l->scratch_a.as<uint64_t>();
        if ((rc = launch_pairs_chunk_offsets(ix->view, dk, ds, de, n, chunk_base, row_rng, st))) return fail(rc);
        uint64_t total = 0;
        cudaError_t e = cudaMemcpyAsync(&total, chunk_base + n_chunks, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error(std::string("pairs total: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
        r->len = total;
        if ((rc = l->out0[0].ensure(std::max<uint64_t>(total, 1) * 4)) || (rc = l->out1[0].ensure(std::max<uint64_t>(total, 1) * 4)))
            return fail(rc);
        if ((rc = pinned_get(ctx, total * 4, &r->blk_probe)) || (rc = pinned_get(ctx, total * 4, &r->blk_build))) return fail(rc);
        if (total) {
            if ((rc = launch_pairs_fill(ix->view, ds, n, chunk_base, row_rng, l->out1[0].as<uint32_t>(), l->out0[0].as<uint32_t>(), st)))
                return fail(rc);
            e = cudaMemcpyAsync(r->blk_probe.p, l->out0[0].p, total * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(r->blk_build.p, l->out1[0].p, total * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { set_error(std::string("pairs D2H: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
        }
        *out = r;
        return BRG_OK;
    }
    uint32_t* counts = l->scratch_a.as<uint32_t>();
    uint64_t* offsets = l->scratch_b.as<uint64_t>();
    if ((rc = launch_pairs_count(ix->view, dk, ds, de, n, counts, st))) return fail(rc);
    if (mode != BRG_PAIRS_INNER) {
        if ((rc = l->scratch_d.ensure(n * 4))) return fail(rc);
        if ((rc = launch_semi_anti_flags(counts, n, mode == BRG_PAIRS_ANTI, l->scratch_d.as<uint32_t>(), st))) return fail(rc);
        counts = l->scratch_d.as<uint32_t>();
    }
    if ((rc = exclusive_scan_u32_to_u64(counts, offsets, n, l->scratch_c.as<uint64_t>(), st))) return fail(rc);
    uint64_t n_use = n, total = 0;
    {
        cudaError_t e = cudaMemcpyAsync(&total, offsets + n, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error(std::string("pairs total: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
    }
    if (max_rows > 0 && total > max_rows) {
        uint64_t* d_cut = reinterpret_cast<uint64_t*>(l->scratch_c.p);
        find_cut_kernel<<<1, 1, 0, st>>>(offsets, n, max_rows, d_cut);
        g_launches.fetch_add(1);
        uint64_t cut[2] = {0, 0};
        cudaError_t e = cudaMemcpyAsync(cut, d_cut, 16, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error(std::string("pairs cut: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
        n_use = cut[0];
        total = cut[1];
    }
    if (consumed) *consumed = n_use;
    r->len = total;
    // device outputs: reuse the lane's output buffers
    if ((rc = l->out0[0].ensure(std::max<uint64_t>(total, 1) * 4))) return fail(rc);
    if (mode == BRG_PAIRS_INNER && (rc = l->out1[0].ensure(std::max<uint64_t>(total, 1) * 4))) return fail(rc);
    if ((rc = pinned_get(ctx, total * 4, &r->blk_probe))) return fail(rc);
    if (mode == BRG_PAIRS_INNER && (rc = pinned_get(ctx, total * 4, &r->blk_build))) return fail(rc);
    if (total) {
        if (mode == BRG_PAIRS_INNER) {
            rc = launch_pairs_emit(ix->view, dk, ds, de, n_use, offsets, l->out1[0].as<uint32_t>(), l->out0[0].as<uint32_t>(), st);
        } else {
            rc = launch_compact_rows(counts, offsets, n_use, l->out0[0].as<uint32_t>(), st);
        }
        if (rc) return fail(rc);
        cudaError_t e = cudaMemcpyAsync(r->blk_probe.p, l->out0[0].p, total * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && mode == BRG_PAIRS_INNER)
            e = cudaMemcpyAsync(r->blk_build.p, l->out1[0].p, total * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error(std::string("pairs D2H: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
    }
    *out = r;
    return BRG_OK;
}

// ---------------------------------------------------------------------------------------------------
// subtract: count -> scan -> emit over the merged intervals of the right side's coverage index
// ---------------------------------------------------------------------------------------------------
extern "C" BRG_API int brg_subtract(brg_ctx* ctx, const brg_index* ix, const uint32_t* key, const int32_t* start, const int32_t* end,
                            uint64_t n, int strict, brg_rows** out) {
    int rc = check_probe(ctx, ix, BRG_WANT_COVERAGE, "brg_subtract");
    if (rc) return rc;
    if (!out) { set_error("brg_subtract: NULL out"); return BRG_ERR_INVALID; }
    *out = nullptr;
    if (n > 0xFFFFFFFFull) { set_error("left row index " + std::to_string(n - 1) + " exceeds UInt32 index capacity"); return BRG_ERR_OVERFLOW; }
    if (n && (!start || !end)) { set_error("brg_subtract: NULL column"); return BRG_ERR_INVALID; }
    BRG_CUDA(cudaSetDevice(ctx->device));
    brg_rows* r = new brg_rows();
    r->ctx = ctx;
    ctx->refs.fetch_add(1);
    if (n == 0) { *out = r; return BRG_OK; }
    LaneGuard g(ctx);
    if ((rc = g.acquire())) { brg_rows_release(r); return rc; }
    Lane* l = g.l;
    cudaStream_t st = l->k;
    auto fail = [&](int code) { cudaStreamSynchronize(st); brg_rows_release(r); return code; };
    if ((rc = upload_probe(l, key, start, end, n))) return fail(rc);
    if ((rc = pairs_scratch(l, n))) return fail(rc);
    if ((rc = l->scratch_b.ensure((n + 1) * 8))) return fail(rc);
    const uint32_t* dk = key ? l->in_key[0].as<uint32_t>() : nullptr;
    const int32_t* ds = l->in_s[0].as<int32_t>();
    const int32_t* de = l->in_e[0].as<int32_t>();
    uint32_t* counts = l->scratch_a.as<uint32_t>();
    uint64_t* offsets = l->scratch_b.as<uint64_t>();
    if ((rc = launch_subtract(ix->view, dk, ds, de, n, strict, counts, nullptr, nullptr, nullptr, nullptr, st))) return fail(rc);
    if ((rc = exclusive_scan_u32_to_u64(counts, offsets, n, l->scratch_c.as<uint64_t>(), st))) return fail(rc);
    uint64_t total = 0;
    {
        cudaError_t e = cudaMemcpyAsync(&total, offsets + n, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error(std::string("subtract total: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
    }
    r->len = total;
    const uint64_t tb = std::max<uint64_t>(total, 1) * 4;
    if ((rc = l->out0[0].ensure(tb)) || (rc = l->out1[0].ensure(tb)) || (rc = l->out2[0].ensure(tb))) return fail(rc);
    if ((rc = pinned_get(ctx, total * 4, &r->blk_probe)) || (rc = pinned_get(ctx, total * 4, &r->blk_ps)) ||
        (rc = pinned_get(ctx, total * 4, &r->blk_pe)))
        return fail(rc);
    if (total) {
        if ((rc = launch_subtract(ix->view, dk, ds, de, n, strict, nullptr, offsets, l->out0[0].as<uint32_t>(), l->out1[0].as<int32_t>(),
                                  l->out2[0].as<int32_t>(), st)))
            return fail(rc);
        cudaError_t e = cudaMemcpyAsync(r->blk_probe.p, l->out0[0].p, total * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r->blk_ps.p, l->out1[0].p, total * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r->blk_pe.p, l->out2[0].p, total * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error(std::string("subtract D2H: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
    }
    *out = r;
    return BRG_OK;
}

// row order of a table sorted by (key, start, end, row): the order in which the sort + sweep table functions walk their
// input (grouped_stream.rs:92-94, 228-237 with contig ids assigned in name order)
extern "C" BRG_API int brg_sort_rows(brg_ctx* ctx, const uint32_t* key, const int32_t* start, const int32_t* end, uint64_t n,
                             uint32_t n_keys, uint32_t* out_row_order) {
    if (!ctx || (n && (!start || !end || !out_row_order))) { set_error("brg_sort_rows: NULL argument"); return BRG_ERR_INVALID; }
    if (n >= 0xFFFFFFFFull) { set_error("brg_sort_rows: row count exceeds UInt32 capacity"); return BRG_ERR_OVERFLOW; }
    if (n_keys == 0 || n_keys == BRG_KEY_ABSENT) { set_error("brg_sort_rows: bad n_keys"); return BRG_ERR_INVALID; }
    if (n == 0) return BRG_OK;
    BRG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->build_stream;
    ScratchScope scratch_scope(st);
    ScratchBuf in[3], setA[4], setB[4], hist, scan;
    int rc;
    for (int i = 0; i < 3; i++) if ((rc = in[i].ensure(n * 4))) return rc;
    for (int i = 0; i < 4; i++) if ((rc = setA[i].ensure(n * 4)) || (rc = setB[i].ensure(n * 4))) return rc;
    if ((rc = hist.ensure(radix_hist_words(n) * 4)) ||
        (rc = scan.ensure((scan_scratch_words(std::max<uint64_t>(radix_hist_words(n), n)) + 16) * 4)))
        return rc;
    if (key) BRG_CUDA(cudaMemcpyAsync(in[0].p, key, n * 4, cudaMemcpyHostToDevice, st));
    BRG_CUDA(cudaMemcpyAsync(in[1].p, start, n * 4, cudaMemcpyHostToDevice, st));
    BRG_CUDA(cudaMemcpyAsync(in[2].p, end, n * 4, cudaMemcpyHostToDevice, st));
    uint32_t* a[4] = {setA[0].as<uint32_t>(), setA[1].as<uint32_t>(), setA[2].as<uint32_t>(), setA[3].as<uint32_t>()};
    uint32_t* b[4] = {setB[0].as<uint32_t>(), setB[1].as<uint32_t>(), setB[2].as<uint32_t>(), setB[3].as<uint32_t>()};
    init_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key ? in[0].as<uint32_t>() : nullptr, in[1].as<int32_t>(),
                                                                  in[2].as<int32_t>(), n, a[0], a[1], a[2], a[3]);
    g_launches.fetch_add(1);
    std::vector<SortPass> passes;
    int key_passes = 0;
    for (uint64_t v = n_keys; v > 0; v >>= 8) key_passes++;
    for (int sh = 0; sh < 32; sh += 8) passes.push_back(SortPass{2, sh, 1, 0xFFFFFFFFu});
    for (int sh = 0; sh < 32; sh += 8) passes.push_back(SortPass{1, sh, 1, 0xFFFFFFFFu});
    for (int p = 0; p < key_passes; p++) passes.push_back(SortPass{0, 8 * p, 0, n_keys});
    int res = 0;
    if ((rc = radix_sort_rows(a, b, 4, n, passes.data(), (int)passes.size(), hist.as<uint32_t>(), hist.bytes / 4, scan.as<uint32_t>(), st, &res)))
        return rc;
    uint32_t** S = res == 0 ? a : b;
    BRG_CUDA(cudaMemcpyAsync(out_row_order, S[3], n * 4, cudaMemcpyDeviceToHost, st));
    BRG_CUDA(cudaStreamSynchronize(st));
    return BRG_OK;
}

// ---------------------------------------------------------------------------------------------------
// merge (the first of the sort + sweep table functions; reuses K1 and K2)
// ---------------------------------------------------------------------------------------------------
struct brg_merged {
    brg_ctx* ctx = nullptr;
    uint64_t len = 0;
    PinnedBlock key{}, start{}, end{}, count{};
};
extern "C" BRG_API uint64_t brg_merged_len(const brg_merged* m) { return m ? m->len : 0; }
extern "C" BRG_API const uint32_t* brg_merged_key(const brg_merged* m) { return m ? (const uint32_t*)m->key.p : nullptr; }
extern "C" BRG_API const int32_t* brg_merged_start(const brg_merged* m) { return m ? (const int32_t*)m->start.p : nullptr; }
extern "C" BRG_API const int32_t* brg_merged_end(const brg_merged* m) { return m ? (const int32_t*)m->end.p : nullptr; }
extern "C" BRG_API const uint32_t* brg_merged_count(const brg_merged* m) { return m ? (const uint32_t*)m->count.p : nullptr; }
extern "C" BRG_API void brg_merged_release(brg_merged* m) {
    if (!m) return;
    pinned_put(m->ctx, m->key); pinned_put(m->ctx, m->start); pinned_put(m->ctx, m->end); pinned_put(m->ctx, m->count);
    ctx_unref(m->ctx);
    delete m;
}

struct brg_gaps {
    brg_ctx* ctx = nullptr;
    uint64_t len = 0;
    PinnedBlock view{}, start{}, end{};
};
extern "C" BRG_API uint64_t brg_gaps_len(const brg_gaps* g) { return g ? g->len : 0; }
extern "C" BRG_API const uint32_t* brg_gaps_view(const brg_gaps* g) { return g ? (const uint32_t*)g->view.p : nullptr; }
extern "C" BRG_API const int64_t* brg_gaps_start(const brg_gaps* g) { return g ? (const int64_t*)g->start.p : nullptr; }
extern "C" BRG_API const int64_t* brg_gaps_end(const brg_gaps* g) { return g ? (const int64_t*)g->end.p : nullptr; }
extern "C" BRG_API void brg_gaps_release(brg_gaps* g) {
    if (!g) return;
    pinned_put(g->ctx, g->view); pinned_put(g->ctx, g->start); pinned_put(g->ctx, g->end);
    ctx_unref(g->ctx);
    delete g;
}
struct ViewArgs {  // complement: the view rows and where the gaps go
    const uint32_t* key;
    const int64_t *start, *end;
    uint64_t n;
    brg_gaps* gaps;
};

static int merge_pipeline(brg_ctx* ctx, const uint32_t* key, const int32_t* start, const int32_t* end, uint64_t n,
                          uint32_t n_keys, int64_t min_dist, int strict, bool cluster, brg_merged** out,
                          const ViewArgs* views = nullptr);
extern "C" BRG_API int brg_complement(brg_ctx* ctx, const uint32_t* key, const int32_t* start, const int32_t* end, uint64_t n,
                              uint32_t n_keys, const uint32_t* view_key, const int64_t* view_start, const int64_t* view_end,
                              uint64_t n_views, int strict, brg_gaps** out) {
    if (!ctx || !out) { set_error("brg_complement: NULL ctx/out"); return BRG_ERR_INVALID; }
    *out = nullptr;
    if (n_views && (!view_key || !view_start || !view_end)) { set_error("brg_complement: NULL view column"); return BRG_ERR_INVALID; }
    if (n_views >= 0xFFFFFFFFull) { set_error("brg_complement: view row count exceeds UInt32 capacity"); return BRG_ERR_OVERFLOW; }
    brg_gaps* g = new brg_gaps();
    g->ctx = ctx;
    ctx->refs.fetch_add(1);
    if (n_views == 0) { *out = g; return BRG_OK; }
    ViewArgs va{view_key, view_start, view_end, n_views, g};
    brg_merged* unused = nullptr;
    const int rc = merge_pipeline(ctx, key, start, end, n, n_keys ? n_keys : 1u, 0, strict, false, &unused, &va);
    if (unused) brg_merged_release(unused);
    if (rc) { brg_gaps_release(g); return rc; }
    *out = g;
    return BRG_OK;
}
extern "C" BRG_API int brg_merge(brg_ctx* ctx, const uint32_t* key, const int32_t* start, const int32_t* end, uint64_t n,
                         uint32_t n_keys, int64_t min_dist, int strict, brg_merged** out) {
    return merge_pipeline(ctx, key, start, end, n, n_keys, min_dist, strict, false, out);
}
extern "C" BRG_API int brg_cluster(brg_ctx* ctx, const uint32_t* key, const int32_t* start, const int32_t* end, uint64_t n,
                           uint32_t n_keys, int64_t min_dist, int strict, brg_merged** out) {
    return merge_pipeline(ctx, key, start, end, n, n_keys, min_dist, strict, true, out);
}
// complement with the merged intervals (m of them, key order) on the device: count -> scan -> emit over the views
static int complement_views(brg_ctx* ctx, const uint32_t* m_key, const int32_t* m_first, const int32_t* m_last, uint64_t m,
                            uint32_t n_keys, const ViewArgs& va, uint32_t* scan_scratch, cudaStream_t st) {
    ScratchScope scratch_scope(st);
    ScratchBuf v_key, v_start, v_end, m_off, g_tmp, G, plan, counts, offsets, scan64, o_view, o_start, o_end;
    int rc;
    const uint64_t nv = va.n;
    if ((rc = v_key.ensure(nv * 4)) || (rc = v_start.ensure(nv * 8)) || (rc = v_end.ensure(nv * 8)) ||
        (rc = m_off.ensure(((uint64_t)n_keys + 1) * 4)) || (rc = g_tmp.ensure(std::max<uint64_t>(m, 1) * 4)) ||
        (rc = G.ensure((m + 1) * 4)) || (rc = plan.ensure(nv * 16)) || (rc = counts.ensure(nv * 4)) ||
        (rc = offsets.ensure((nv + 1) * 8)) || (rc = scan64.ensure((scan_scratch_words(nv) + 16) * 8)))
        return rc;
    BRG_CUDA(cudaMemcpyAsync(v_key.p, va.key, nv * 4, cudaMemcpyHostToDevice, st));
    BRG_CUDA(cudaMemcpyAsync(v_start.p, va.start, nv * 8, cudaMemcpyHostToDevice, st));
    BRG_CUDA(cudaMemcpyAsync(v_end.p, va.end, nv * 8, cudaMemcpyHostToDevice, st));
    BRG_CUDA(cudaMemsetAsync(G.p, 0, (m + 1) * 4, st));
    if ((rc = complement_plan(m_key, m_first, m_last, m, n_keys, v_key.as<uint32_t>(), v_start.as<long long>(), v_end.as<long long>(),
                              nv, m_off.as<uint32_t>(), g_tmp.as<uint32_t>(), G.as<uint32_t>(), scan_scratch, plan.as<uint4>(),
                              counts.as<uint32_t>(), st)))
        return rc;
    if ((rc = exclusive_scan_u32_to_u64(counts.as<uint32_t>(), offsets.as<uint64_t>(), nv, scan64.as<uint64_t>(), st))) return rc;
    uint64_t total = 0;
    BRG_CUDA(cudaMemcpyAsync(&total, offsets.as<uint64_t>() + nv, 8, cudaMemcpyDeviceToHost, st));
    BRG_CUDA(cudaStreamSynchronize(st));
    brg_gaps* g = va.gaps;
    g->len = total;
    if (total == 0) return 0;
    if ((rc = o_view.ensure(total * 4)) || (rc = o_start.ensure(total * 8)) || (rc = o_end.ensure(total * 8))) return rc;
    if ((rc = pinned_get(ctx, total * 4, &g->view)) || (rc = pinned_get(ctx, total * 8, &g->start)) ||
        (rc = pinned_get(ctx, total * 8, &g->end)))
        return rc;
    if ((rc = complement_emit(v_start.as<long long>(), v_end.as<long long>(), nv, offsets.as<uint64_t>(), m_first, m_last,
                              G.as<uint32_t>(), plan.as<uint4>(), total, o_view.as<uint32_t>(), o_start.as<long long>(),
                              o_end.as<long long>(), st)))
        return rc;
    BRG_CUDA(cudaMemcpyAsync(g->view.p, o_view.p, total * 4, cudaMemcpyDeviceToHost, st));
    BRG_CUDA(cudaMemcpyAsync(g->start.p, o_start.p, total * 8, cudaMemcpyDeviceToHost, st));
    BRG_CUDA(cudaMemcpyAsync(g->end.p, o_end.p, total * 8, cudaMemcpyDeviceToHost, st));
    BRG_CUDA(cudaStreamSynchronize(st));
    return 0;
}

static int merge_pipeline(brg_ctx* ctx, const uint32_t* key, const int32_t* start, const int32_t* end, uint64_t n,
                          uint32_t n_keys, int64_t min_dist, int strict, bool cluster, brg_merged** out, const ViewArgs* views) {
    if (!ctx || !out) { set_error("brg_merge: NULL ctx/out"); return BRG_ERR_INVALID; }
    *out = nullptr;
    if (n > 0 && (!start || !end)) { set_error("brg_merge: NULL column"); return BRG_ERR_INVALID; }
    if (n >= 0xFFFFFFFFull) { set_error("brg_merge: row count exceeds UInt32 capacity"); return BRG_ERR_OVERFLOW; }
    if (n_keys == 0 || n_keys == BRG_KEY_ABSENT) { set_error("brg_merge: bad n_keys"); return BRG_ERR_INVALID; }
    BRG_CUDA(cudaSetDevice(ctx->device));
    brg_merged* r = new brg_merged();
    r->ctx = ctx;
    ctx->refs.fetch_add(1);
    cudaStream_t st = ctx->build_stream;
    ScratchScope scratch_scope(st);
    if (n == 0) {
        *out = r;
        if (!views) return BRG_OK;
        ScratchBuf scan0;  // no input rows: no merged intervals, every view is answered from its own bounds
        int rc0 = scan0.ensure(64);
        if (!rc0) rc0 = complement_views(ctx, nullptr, nullptr, nullptr, 0, n_keys, *views, scan0.as<uint32_t>(), st);
        if (rc0) cudaStreamSynchronize(st);
        return rc0;
    }
    ScratchBuf in[3], setA[4], setB[4], hist, scan, small, pme, flags, midx, m_key, m_first, m_last, m_row0, m_count;
    int rc;
    auto fail = [&](int code) { cudaStreamSynchronize(st); brg_merged_release(r); return code; };
    for (int i = 0; i < 3; i++) if ((rc = in[i].ensure(n * 4))) return fail(rc);
    for (int i = 0; i < 4; i++) if ((rc = setA[i].ensure(n * 4)) || (rc = setB[i].ensure(n * 4))) return fail(rc);
    if ((rc = hist.ensure(radix_hist_words(n) * 4)) || (rc = small.ensure(4096)) ||
        (rc = scan.ensure((scan_scratch_words(std::max<uint64_t>(radix_hist_words(n), n)) + 16) * 4)))
        return fail(rc);
    if (key) BRG_CUDA(cudaMemcpyAsync(in[0].p, key, n * 4, cudaMemcpyHostToDevice, st));
    BRG_CUDA(cudaMemcpyAsync(in[1].p, start, n * 4, cudaMemcpyHostToDevice, st));
    BRG_CUDA(cudaMemcpyAsync(in[2].p, end, n * 4, cudaMemcpyHostToDevice, st));
    uint32_t* a[4] = {setA[0].as<uint32_t>(), setA[1].as<uint32_t>(), setA[2].as<uint32_t>(), setA[3].as<uint32_t>()};
    uint32_t* b[4] = {setB[0].as<uint32_t>(), setB[1].as<uint32_t>(), setB[2].as<uint32_t>(), setB[3].as<uint32_t>()};
    init_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key ? in[0].as<uint32_t>() : nullptr, in[1].as<int32_t>(),
                                                                  in[2].as<int32_t>(), n, a[0], a[1], a[2], a[3]);
    g_launches.fetch_add(1);
    // intervals.sort_unstable() per contig (grouped_stream.rs:92-94): rows ordered by (key, start, end)
    std::vector<SortPass> passes;
    int key_passes = 0;
    for (uint64_t v = n_keys; v > 0; v >>= 8) key_passes++;
    for (int sh = 0; sh < 32; sh += 8) passes.push_back(SortPass{2, sh, 1, 0xFFFFFFFFu});
    for (int sh = 0; sh < 32; sh += 8) passes.push_back(SortPass{1, sh, 1, 0xFFFFFFFFu});
    for (int p = 0; p < key_passes; p++) passes.push_back(SortPass{0, 8 * p, 0, n_keys});
    int res = 0;
    // (cluster carries the row ids along: ties keep input order, like the reference's (start, end, row) tuples)
    if ((rc = radix_sort_rows(a, b, cluster ? 4 : 3, n, passes.data(), (int)passes.size(), hist.as<uint32_t>(), hist.bytes / 4, scan.as<uint32_t>(), st, &res)))
        return fail(rc);
    uint32_t** S = res == 0 ? a : b;
    uint32_t* d_scalar = small.as<uint32_t>();
    // rows with start > end break the equivalence "run end == running max": the sequential sweep has no parallel form then
    if ((rc = flag_inverted_rows((const int32_t*)S[1], (const int32_t*)S[2], n, d_scalar + 8, st))) return fail(rc);
    if ((rc = pme.ensure(n * 4)) || (rc = flags.ensure(n * 4)) || (rc = midx.ensure(n * 4)) || (rc = m_key.ensure(n * 4)) ||
        (rc = m_first.ensure(n * 4)) || (rc = m_last.ensure(n * 4)) || (rc = m_row0.ensure(n * 4)) || (rc = m_count.ensure(n * 4)))
        return fail(rc);
    if ((rc = segmented_running_max(S[0], (const int32_t*)S[2], pme.as<int32_t>(), n, scan.as<int32_t>(), st))) return fail(rc);
    if ((rc = merge_runs(S[0], (const int32_t*)S[1], pme.as<int32_t>(), n, (long long)min_dist, strict, flags.as<uint32_t>(),
                         midx.as<uint32_t>(), scan.as<uint32_t>(), m_key.as<uint32_t>(), m_first.as<int32_t>(),
                         m_last.as<int32_t>(), m_row0.as<uint32_t>(), d_scalar, st)))
        return fail(rc);
    uint32_t host_vals[9] = {0};
    {
        cudaError_t e = cudaMemcpyAsync(host_vals, d_scalar, sizeof(host_vals), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error(std::string("merge: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
    }
    if (host_vals[8]) { set_error("merge: intervals with start > end are not supported on the GPU path"); return fail(BRG_ERR_UNSUPPORTED); }
    uint64_t m = host_vals[0];
    if (views) {
        if ((rc = complement_views(ctx, m_key.as<uint32_t>(), m_first.as<int32_t>(), m_last.as<int32_t>(), m, n_keys, *views,
                                   scan.as<uint32_t>(), st)))
            return fail(rc);
        *out = r;
        return BRG_OK;
    }
    if (cluster) {
        // one result row per input row, in (key, start, end, row) order:
        //   key -> original row id, start / end -> extent of the row's run, count -> run id (global, key order)
        // flags / midx stay as they are; the run heads move out of the way first (m_count, m_row0, pme as outputs)
        ScratchBuf c_start, c_end;
        if ((rc = c_start.ensure(n * 4)) || (rc = c_end.ensure(n * 4))) return fail(rc);
        if ((rc = cluster_rows(flags.as<uint32_t>(), midx.as<uint32_t>(), m_first.as<int32_t>(), m_last.as<int32_t>(), n,
                               m_count.as<uint32_t>(), c_start.as<int32_t>(), c_end.as<int32_t>(), st)))
            return fail(rc);
        cudaError_t ce = cudaMemcpyAsync(m_key.p, S[3], n * 4, cudaMemcpyDeviceToDevice, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(m_first.p, c_start.p, n * 4, cudaMemcpyDeviceToDevice, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(m_last.p, c_end.p, n * 4, cudaMemcpyDeviceToDevice, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { set_error(std::string("cluster: ") + cudaGetErrorString(ce)); return fail(BRG_ERR_CUDA); }
        m = n;
    } else if ((rc = merge_counts(m_row0.as<uint32_t>(), m, n, m_count.as<uint32_t>(), st))) {
        return fail(rc);
    }
    r->len = m;
    if ((rc = pinned_get(ctx, m * 4, &r->key)) || (rc = pinned_get(ctx, m * 4, &r->start)) || (rc = pinned_get(ctx, m * 4, &r->end)) ||
        (rc = pinned_get(ctx, m * 4, &r->count)))
        return fail(rc);
    cudaError_t e = cudaMemcpyAsync(r->key.p, m_key.p, m * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->start.p, m_first.p, m * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->end.p, m_last.p, m * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->count.p, m_count.p, m * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error(std::string("merge D2H: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
    *out = r;
    return BRG_OK;
}

extern "C" BRG_API int brg_nearest(brg_ctx* ctx, const brg_index* ix, const uint32_t* key, const int32_t* start,
                           const int32_t* end, uint64_t n, uint32_t k, int include_overlaps, int compute_distance,
                           int strict, brg_rows** out) {
    int rc = check_probe(ctx, ix, BRG_WANT_NEAREST, "brg_nearest");
    if (rc) return rc;
    if (!out) { set_error("brg_nearest: NULL out"); return BRG_ERR_INVALID; }
    *out = nullptr;
    if (k == 0) { set_error("nearest: k must be >= 1"); return BRG_ERR_INVALID; }  // table_function.rs:191-214
    if (n > 0xFFFFFFFFull) { set_error("right row index " + std::to_string(n - 1) + " exceeds UInt32 index capacity"); return BRG_ERR_OVERFLOW; }
    if (n && (!start || !end)) { set_error("brg_nearest: NULL column"); return BRG_ERR_INVALID; }
    BRG_CUDA(cudaSetDevice(ctx->device));
    brg_rows* r = new brg_rows();
    r->ctx = ctx;
    ctx->refs.fetch_add(1);
    if (n == 0) { *out = r; return BRG_OK; }
    LaneGuard g(ctx);
    if ((rc = g.acquire())) { brg_rows_release(r); return rc; }
    Lane* l = g.l;
    cudaStream_t st = l->k;
    auto fail = [&](int code) { cudaStreamSynchronize(st); brg_rows_release(r); return code; };
    if ((rc = upload_probe(l, key, start, end, n))) return fail(rc);
    if ((rc = pairs_scratch(l, n))) return fail(rc);
    if ((rc = l->scratch_b.ensure((n + 1) * 8))) return fail(rc);
    const uint32_t* dk = key ? l->in_key[0].as<uint32_t>() : nullptr;
    const int32_t* ds = l->in_s[0].as<int32_t>();
    const int32_t* de = l->in_e[0].as<int32_t>();
    uint32_t* counts = l->scratch_a.as<uint32_t>();
    uint64_t* offsets = l->scratch_b.as<uint64_t>();
    if ((rc = launch_nearest_k_count(ix->view, dk, ds, de, n, k, include_overlaps, strict, counts, st))) return fail(rc);
    if ((rc = exclusive_scan_u32_to_u64(counts, offsets, n, l->scratch_c.as<uint64_t>(), st))) return fail(rc);
    uint64_t total = 0;
    {
        cudaError_t e = cudaMemcpyAsync(&total, offsets + n, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error(std::string("nearest total: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
    }
    r->len = total;
    const uint64_t bit_words = (total + 31) / 32;
    if ((rc = l->out0[0].ensure(total * 4)) || (rc = l->out1[0].ensure(total * 4)) || (rc = l->scratch_d.ensure(total * 4)) ||
        (rc = l->out0[1].ensure(bit_words * 4 + 4)))
        return fail(rc);
    if (compute_distance && (rc = l->out2[0].ensure(total * 8))) return fail(rc);
    if ((rc = pinned_get(ctx, total * 4, &r->blk_build)) || (rc = pinned_get(ctx, total * 4, &r->blk_probe)) ||
        (rc = pinned_get(ctx, bit_words * 4 + 4, &r->blk_bits)))
        return fail(rc);
    if (compute_distance && (rc = pinned_get(ctx, total * 8, &r->blk_dist))) return fail(rc);
    int64_t* d_dist = compute_distance ? l->out2[0].as<int64_t>() : nullptr;
    rc = launch_nearest_k_emit(ix->view, dk, ds, de, n, k, include_overlaps, strict, offsets, l->out0[0].as<uint32_t>(),
                               l->scratch_d.as<uint32_t>(), l->out1[0].as<uint32_t>(), d_dist, st);
    if (rc) return fail(rc);
    pack_bits_kernel<<<(unsigned)((((total + 31) & ~31ull) + 255) / 256), 256, 0, st>>>(l->scratch_d.as<uint32_t>(), total,
                                                                                         l->out0[1].as<uint32_t>());
    g_launches.fetch_add(1);
    cudaError_t e = cudaMemcpyAsync(r->blk_build.p, l->out0[0].p, total * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->blk_probe.p, l->out1[0].p, total * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r->blk_bits.p, l->out0[1].p, bit_words * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && compute_distance) e = cudaMemcpyAsync(r->blk_dist.p, d_dist, total * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error(std::string("nearest D2H: ") + cudaGetErrorString(e)); return fail(BRG_ERR_CUDA); }
    *out = r;
    return BRG_OK;
}
