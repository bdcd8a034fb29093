// brg_win.cuh -- a chunk's WINDOW of a sorted build array in shared memory, and the searches over it (sm_100a).
// Shared by the one-launch pair join (brg_join.cu) and the windowed nearest kernel (brg_probe.cu).
//
// A chunk of coordinate-sorted probe rows of one key touches a short run of each sorted build array.  The run starts at the first
// row of the rank-directory bin of (the chunk's smallest start) - 1 -- ONE dependent load instead of a rank query, at the price of a
// few extra rows -- and is copied into shared memory by TMA (16-byte aligned source: up to 3 rows of slack in front, `wlo` of
// which may belong to the previous key and are skipped).  All ranks of the chunk then come from branch-free binary searches over
// the window with a warp-uniform trip count and NO bound check per step: behind the window's last row stand INT32_MAX sentinels as
// far as a search can reach.
#pragma once
#include "brg_internal.cuh"

namespace brg {

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t bit_floor_u32(uint32_t x) { return x ? 1u << (31 - __clz(x)) : 0u; }
__host__ __device__ constexpr int win_bit_floor(int x) { int b = 1; while (2 * b <= x) b *= 2; return b; }
// sentinel rows behind a window of `cap` rows: a search whose first step is S = bit_floor(cap) reads up to row 2 S + 1
__host__ __device__ constexpr int win_pad(int cap) { return 2 * win_bit_floor(cap) + 8 - cap; }

// the first row of key k whose directory bin is bin(q_first - 1): every earlier row of the key is < q_first
__device__ __forceinline__ uint32_t win_begin(const DirView& d, const KeyDir& k, int32_t q_first) {
    if (k.nb == 0 || q_first == INT32_MIN || q_first - 1 < k.lo) return k.seg_begin;
    const uint32_t b = ((uint32_t)(q_first - 1) - (uint32_t)k.lo) >> d.shift;
    if (b >= k.nb) return k.seg_end;
    return __ldg(d.D + k.base + b);
}

// INT32_MAX behind the window's last row, as far as win_ranks<.., CAP> can reach (w has room for CAP + win_pad(CAP) rows)
template <int CAP>
__device__ __forceinline__ void win_sentinels(int32_t* w, uint32_t nw, uint32_t lane) {
    constexpr uint32_t END = 2u * (uint32_t)win_bit_floor(CAP) + 4u;
    for (uint32_t x = nw + lane; x < END; x += 32) w[x] = INT32_MAX;  // (a full window: one to three stores)
}
// The same in two parts for kernels whose window arrays are not reused for anything else: the rows behind a FULL window are
// written once per kernel (a TMA copy of at most CAP rows never touches them), and a chunk whose window is cut short by the end
// of its key -- the rare case -- fills the rest.
template <int CAP>
__device__ __forceinline__ void win_sentinels_init(int32_t* w, uint32_t lane) { win_sentinels<CAP>(w, (uint32_t)CAP, lane); }
template <int CAP>
__device__ __forceinline__ void win_sentinels_short(int32_t* w, uint32_t nw, uint32_t lane) {
    if (nw < (uint32_t)CAP)
        for (uint32_t x = nw + lane; x < (uint32_t)CAP; x += 32) w[x] = INT32_MAX;
}

// ranks of NQ queries in the window rows w[wlo .. nw), nw <= CAP: out[j] = #(w[0 .. nw) <= q[j]) (LE) or #(... < q[j]), counting the
// wlo rows in front as smaller.  NQ independent searches in lock step, ALWAYS log2(bit_floor(CAP)) + 1 steps, fully unrolled: a step
// is a load at a constant offset, a compare and a predicated add (a window is almost always full: a trip count that follows nw
// saves nothing and costs a loop counter, an add and a bound per step).  A sentinel passes only `<= INT32_MAX`, so results are
// clamped once at the end.  Positions are byte offsets of w[pos - 1]; w[-1] must be readable.  (bitwise logic only: with `&&` and a
// bound check per step the compiler emitted a divergent branch per step -- 22 % of the first join kernel's instructions were
// BRA / BSSY / BSYNC.)
// Lanes should hold CONSECUTIVE sorted queries (row = lane + 32 j): their positions in the window are then about one row apart --
// broadcasts and consecutive banks instead of 4-way conflicts.
template <bool LE, int NQ, int CAP>
__device__ __forceinline__ void win_ranks(const int32_t* w, uint32_t wlo, uint32_t nw, const int32_t (&q)[NQ], uint32_t (&out)[NQ]) {
    const char* wb = reinterpret_cast<const char*>(w) - 4;
    uint32_t ob[NQ];
#pragma unroll
    for (int j = 0; j < NQ; j++) ob[j] = wlo * 4u;
#pragma unroll
    for (uint32_t step = 4u * (uint32_t)win_bit_floor(CAP); step >= 4u; step >>= 1) {
#pragma unroll
        for (int j = 0; j < NQ; j++) {
            const int32_t v = *reinterpret_cast<const int32_t*>(wb + ob[j] + step);
            if (LE ? v <= q[j] : v < q[j]) ob[j] += step;
        }
    }
#pragma unroll
    for (int j = 0; j < NQ; j++) out[j] = min(ob[j] >> 2, nw);
}
#endif

}  // namespace brg
