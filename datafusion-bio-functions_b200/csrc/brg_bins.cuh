// brg_bins.cuh -- the 32-byte bin record shared by the bins tables (count bins, coverage bins, nearest / pair bins).
//
// Plain C++ on purpose (no CUDA intrinsics): nvcc compiles it into the build and probe kernels, g++ compiles it into
// tests/test_bins_record.py's host harness, which checks the encoder against the decoder exhaustively on the CPU.
//
// One record describes one coordinate bin [thr, thr + 2^shift) of a PAIR of sorted arrays A, B that share their key segments:
//   w[0]        cs    = global index of the first element of A that is >= thr
//   w[1] 0..15  depth = cs - (global index of the first element of B that is >= thr)
//   w[1] 16..23 na    = elements of A inside the bin
//   w[1] 24..31 nb    = elements of B inside the bin; 0xFF marks a record that does not fit (the row searches the arrays)
//   w[2..7]     the in-bin offsets of the na elements of A, then of the nb elements of B, each group ascending:
//               format 0: 12 slots of 16 bits (bins up to 65536 wide)
//               format 1: 18 slots of 10 bits, three per word (bins up to 1024 wide) -- half the bytes per element (a table over
//                         10M build rows is 97 MB instead of 194 MB).  Slot i lives in word 2 + i % 6, field i / 6 (bits
//                         10 * (i / 6) .. + 9): the decoder compares a whole word against the query offset with two subtractions
//                         (11-bit lanes behind guard bits) and shifts word j's three result bits to bit positions j, 10 + j,
//                         20 + j, so that slot i answers at bit 10 * (i / 6) + i % 6 -- increasing with i: "the slots below n" is
//                         one mask, and both rank queries of a row cost about 60 instructions where the slot-by-slot loop takes 110
#ifndef BRG_BINS_CUH
#define BRG_BINS_CUH
#include <stdint.h>

#ifdef __CUDACC__
#define BRG_HD __host__ __device__ __forceinline__
#else
#define BRG_HD inline
#endif

namespace brg {

constexpr uint32_t CB_OVERFLOW = 0xFFu;
constexpr uint32_t CB_SLOTS_WIDE = 12;    // format 0
constexpr uint32_t CB_SLOTS_PACKED = 18;  // format 1
constexpr uint32_t CB_SHIFT_WIDE = 16;    // widest bin (log2) whose offsets fit a slot
constexpr uint32_t CB_SHIFT_PACKED = 10;

BRG_HD uint32_t cb_slots(uint32_t fmt) { return fmt ? CB_SLOTS_PACKED : CB_SLOTS_WIDE; }
BRG_HD uint32_t cb_max_shift(uint32_t fmt) { return fmt ? CB_SHIFT_PACKED : CB_SHIFT_WIDE; }

template <int FMT>
BRG_HD uint32_t cb_slot_get(const uint32_t (&w)[8], uint32_t i) {
    if (FMT) return (w[2 + i % 6] >> ((i / 6) * 10u)) & 0x3FFu;
    return (w[2 + (i >> 1)] >> ((i & 1u) * 16u)) & 0xFFFFu;
}
template <int FMT>
BRG_HD void cb_slot_or(uint32_t (&w)[8], uint32_t i, uint32_t v) {
    if (FMT) w[2 + i % 6] |= (v & 0x3FFu) << ((i / 6) * 10u);
    else w[2 + (i >> 1)] |= (v & 0xFFFFu) << ((i & 1u) * 16u);
}
BRG_HD uint32_t cb_popc(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
// Packed format: one bit per slot whose offset is <= t; slot i answers at bit 10 * (i / 6) + i % 6.
// Per word: fields 0 and 2 in one subtraction, field 1 in another; each lane is (1024 + t) - v >= 1, so no borrow leaves a lane
// and the guard bit (10 / 20 / 30) survives exactly where v <= t.
BRG_HD uint32_t cb_packed_le_mask(const uint32_t (&w)[8], uint32_t t) {
    const uint32_t te = (t | (t << 20)) | 0x40000400u, to = (t << 10) | 0x00100000u;
    uint32_t m = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (uint32_t j = 0; j < 6; j++) {
        const uint32_t x = w[2 + j];
        const uint32_t g = ((te - (x & 0x3FF003FFu)) & 0x40000400u) | ((to - (x & 0x000FFC00u)) & 0x00100000u);
        m |= g >> (10u - j);
    }
    return m;
}
// the bits of the packed slots below n (n <= 18) in that mask: n / 6 == (n * 43) >> 8 on that range
BRG_HD uint32_t cb_packed_below(uint32_t n) { return (1u << (n + 4u * ((n * 43u) >> 8))) - 1u; }

BRG_HD bool cb_overflowed(const uint32_t (&w)[8]) { return (w[1] >> 24) == CB_OVERFLOW; }
BRG_HD uint32_t cb_first_a(const uint32_t (&w)[8]) { return w[0]; }
BRG_HD uint32_t cb_first_b(const uint32_t (&w)[8]) { return w[0] - (w[1] & 0xFFFFu); }

// header of a record; returns false (and marks the record) when it does not fit the format.  cs / cb = global index of the first
// element of A / B that is >= the bin's first coordinate, na / nb = elements inside the bin.
BRG_HD bool cb_encode_header(uint32_t (&w)[8], uint32_t fmt, uint32_t cs, uint32_t cb, uint32_t na, uint32_t nb) {
    w[0] = cs;
    for (int i = 1; i < 8; i++) w[i] = 0u;
    // (cb > cs: more of B than of A lies before the bin -- only with rows whose second coordinate is below their first)
    if (na + nb > cb_slots(fmt) || cb > cs || cs - cb >= 0xFFFFu) {
        w[1] = CB_OVERFLOW << 24;
        return false;
    }
    w[1] = (cs - cb) | (na << 16) | (nb << 24);
    return true;
}
// slot i (A's elements first, then B's) = offset of the element from the bin's first coordinate
BRG_HD void cb_encode_slot(uint32_t (&w)[8], uint32_t fmt, uint32_t i, uint32_t off) {
    if (fmt) cb_slot_or<1>(w, i, off);
    else cb_slot_or<0>(w, i, off);
}

// #(elements of the bin with offset <= off) among A's (SECOND = false) or B's (true) of a record that did not overflow
template <int FMT, bool SECOND>
BRG_HD uint32_t cb_count_le_fmt(const uint32_t (&w)[8], uint32_t off) {
    const uint32_t na = (w[1] >> 16) & 0xFFu, nb = w[1] >> 24;
    const uint32_t first = SECOND ? na : 0u, last = SECOND ? na + nb : na;
    uint32_t c = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (uint32_t i = 0; i < (FMT ? CB_SLOTS_PACKED : CB_SLOTS_WIDE); i++) {
        const uint32_t v = cb_slot_get<FMT>(w, i);
        c += ((!SECOND || i >= first) && i < last && v <= off) ? 1u : 0u;
    }
    return c;
}
// the packed format without the slot-by-slot loop (cb_count_le_fmt<1, .> is kept as the plain statement the harness compares with)
template <bool SECOND>
BRG_HD uint32_t cb_count_le_packed(const uint32_t (&w)[8], uint32_t off) {
    const uint32_t na = (w[1] >> 16) & 0xFFu, nb = w[1] >> 24;
    const uint32_t m = cb_packed_le_mask(w, off);
    if (!SECOND) return cb_popc(m & cb_packed_below(na));
    return cb_popc(m & cb_packed_below(na + nb) & ~cb_packed_below(na));
}
template <bool SECOND>
BRG_HD uint32_t cb_count_le(const uint32_t (&w)[8], uint32_t fmt, uint32_t off) {
    return fmt ? cb_count_le_packed<SECOND>(w, off) : cb_count_le_fmt<0, SECOND>(w, off);
}

}  // namespace brg
#endif
