// Host harness for csrc/brg_bins.cuh (tests/test_bins_record.py compiles and runs it with g++; no GPU).
// Builds the records of one key exactly as cb_fill_kernel does (brg_scan.cu) and answers both rank queries of a row as
// bins_pair_ranks does (brg_internal.cuh), for every coordinate around the key's range, against a brute-force count.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "brg_bins.cuh"

using namespace brg;

struct Rec { uint32_t w[8]; };

static int run_case(uint32_t seed, uint32_t fmt, uint32_t shift, uint32_t n, uint32_t span, bool allow_second_below, uint64_t* overflowed, uint64_t* answered) {
    std::mt19937 rng(seed);
    const int32_t origin = (int32_t)(rng() % 2000) - 1000;
    std::vector<int32_t> A(n), B(n);
    for (uint32_t i = 0; i < n; i++) {
        A[i] = origin + (int32_t)(rng() % span);
        const int32_t len = (int32_t)(rng() % 64);
        B[i] = allow_second_below && (rng() % 8 == 0) ? A[i] - len : A[i] + len;
    }
    std::sort(A.begin(), A.end());
    std::sort(B.begin(), B.end());
    const uint32_t seg_begin = 1000, seg_end = seg_begin + n;  // global indices of the key's segment
    const int32_t lo = std::min(A.front(), B.front()), hi = std::max(A.back(), B.back());
    const uint32_t nb = (((uint32_t)hi - (uint32_t)lo) >> shift) + 1;
    std::vector<Rec> rec(nb + 1);
    for (uint32_t j = 0; j <= nb; j++) {
        uint32_t (&w)[8] = rec[j].w;
        for (int i = 0; i < 8; i++) w[i] = 0;
        w[0] = seg_end;  // the key's sentinel record
        if (j == nb) break;
        const int64_t thr = (int64_t)lo + ((int64_t)j << shift), thr1 = thr + (1ll << shift);
        const uint32_t cs = seg_begin + (uint32_t)(std::lower_bound(A.begin(), A.end(), thr, [](int32_t a, int64_t t) { return a < t; }) - A.begin());
        const uint32_t ce = seg_begin + (uint32_t)(std::lower_bound(B.begin(), B.end(), thr, [](int32_t a, int64_t t) { return a < t; }) - B.begin());
        const uint32_t cs1 = seg_begin + (uint32_t)(std::lower_bound(A.begin(), A.end(), thr1, [](int32_t a, int64_t t) { return a < t; }) - A.begin());
        const uint32_t ce1 = seg_begin + (uint32_t)(std::lower_bound(B.begin(), B.end(), thr1, [](int32_t a, int64_t t) { return a < t; }) - B.begin());
        const uint32_t ns = cs1 - cs, ne = ce1 - ce;
        if (cb_encode_header(w, fmt, cs, ce, ns, ne)) {
            for (uint32_t i = 0; i < ns + ne; i++) {
                const int64_t v = (int64_t)(i < ns ? A[cs - seg_begin + i] : B[ce - seg_begin + (i - ns)]) - thr;
                if (v < 0 || v >= (1ll << shift)) { std::printf("offset out of range\n"); return 1; }
                cb_encode_slot(w, fmt, i, (uint32_t)v);
            }
        } else {
            (*overflowed)++;
        }
    }
    // queries: ra = seg_begin + #(A <= qa), rb = seg_begin + #(B < qb)
    const uint32_t msk = (1u << shift) - 1u;
    for (int64_t q64 = (int64_t)lo - 70; q64 <= (int64_t)hi + 70; q64++) {
        const int32_t qa = (int32_t)q64, qb = (int32_t)q64;
        const bool a_in = qa >= lo, b_in = qb > lo;
        const uint32_t da = (uint32_t)qa - (uint32_t)lo, db = (uint32_t)qb - 1u - (uint32_t)lo;
        const uint32_t ba = da >> shift, bb = db >> shift;
        uint32_t ra = a_in ? seg_end : seg_begin, rb = b_in ? seg_end : seg_begin;
        bool miss_a = false, miss_b = false;
        if (a_in && ba < nb) {
            const uint32_t (&w)[8] = rec[ba].w;
            if (cb_overflowed(w)) miss_a = true;
            else ra = cb_first_a(w) + cb_count_le<false>(w, fmt, da & msk);
            if (!miss_a && fmt && cb_count_le<false>(w, fmt, da & msk) != cb_count_le_fmt<1, false>(w, da & msk)) { std::printf("packed count differs from the slot loop (A)\n"); return 1; }
        }
        if (b_in && bb < nb) {
            const uint32_t (&w)[8] = rec[bb].w;
            if (cb_overflowed(w)) miss_b = true;
            else rb = cb_first_b(w) + cb_count_le<true>(w, fmt, db & msk);
            if (!miss_b && fmt && cb_count_le<true>(w, fmt, db & msk) != cb_count_le_fmt<1, true>(w, db & msk)) { std::printf("packed count differs from the slot loop (B)\n"); return 1; }
        }
        *answered += (a_in && ba < nb && !miss_a) + (b_in && bb < nb && !miss_b);
        const uint32_t want_a = seg_begin + (uint32_t)(std::upper_bound(A.begin(), A.end(), qa) - A.begin());
        const uint32_t want_b = seg_begin + (uint32_t)(std::lower_bound(B.begin(), B.end(), qb) - B.begin());
        if ((!miss_a && ra != want_a) || (!miss_b && rb != want_b)) {
            std::printf("MISMATCH seed %u fmt %u shift %u q %lld: ra %u want %u, rb %u want %u\n", seed, fmt, shift, (long long)q64, ra, want_a, rb, want_b);
            return 1;
        }
    }
    return 0;
}

int main() {
    uint64_t cases = 0, overflowed = 0, answered = 0;
    for (uint32_t fmt = 0; fmt < 2; fmt++)
        for (uint32_t shift = 0; shift <= cb_max_shift(fmt); shift++)
            for (uint32_t seed = 0; seed < 6; seed++) {
                // density tuned so that some records overflow and most do not
                const uint32_t span = std::max<uint32_t>(64u, (200u << shift) / (seed + 1));
                const uint32_t n = 300 + 50 * seed;
                if (run_case(1000 * fmt + 10 * shift + seed, fmt, shift, n, std::min<uint32_t>(span, 3000000u), seed % 2 == 1, &overflowed, &answered)) return 1;
                cases++;
            }
    // the limits of the formats
    if (cb_slots(0) != 12 || cb_slots(1) != 18 || cb_max_shift(0) != 16 || cb_max_shift(1) != 10) { std::printf("format constants\n"); return 1; }
    uint32_t w[8];
    if (cb_encode_header(w, 0, 10, 5, 7, 6) || !cb_overflowed(w)) { std::printf("13 elements must overflow format 0\n"); return 1; }
    if (!cb_encode_header(w, 1, 10, 5, 9, 9) || cb_overflowed(w)) { std::printf("18 elements must fit format 1\n"); return 1; }
    if (cb_encode_header(w, 1, 10, 5, 10, 9)) { std::printf("19 elements must overflow format 1\n"); return 1; }
    if (cb_encode_header(w, 1, 5, 10, 1, 1)) { std::printf("negative depth must overflow\n"); return 1; }
    if (cb_encode_header(w, 0, 70000, 4, 1, 1)) { std::printf("depth >= 65535 must overflow\n"); return 1; }
    for (uint32_t i = 0; i < 18; i++) {  // every packed slot holds 10 bits without touching its neighbours
        cb_encode_header(w, 1, 0, 0, 9, 9);
        cb_encode_slot(w, 1, i, 1023);
        for (uint32_t j = 0; j < 18; j++)
            if (cb_slot_get<1>(w, j) != (j == i ? 1023u : 0u)) { std::printf("packed slot %u leaks into %u\n", i, j); return 1; }
    }
    if (answered < 1000000 || overflowed < 100) { std::printf("vacuous: %llu ranks answered from records, %llu overflowed records\n", (unsigned long long)answered, (unsigned long long)overflowed); return 1; }
    {   // the packed decoder against the slot loop on arbitrary slot contents (not only ascending ones), every split, many offsets
        std::mt19937 rng(7);
        for (int it = 0; it < 20000; it++) {
            const uint32_t na = rng() % 19, nb = rng() % (19 - na);
            cb_encode_header(w, 1, 100, 50, na, nb);
            for (uint32_t i = 0; i < na + nb; i++) cb_encode_slot(w, 1, i, (it % 3 == 0) ? (rng() % 4) * 341u : rng() % 1024);
            for (int q = 0; q < 24; q++) {
                const uint32_t off = q < 4 ? (uint32_t[]){0u, 1u, 1022u, 1023u}[q] : rng() % 1024;
                if (cb_count_le<false>(w, 1, off) != cb_count_le_fmt<1, false>(w, off) || cb_count_le<true>(w, 1, off) != cb_count_le_fmt<1, true>(w, off)) {
                    std::printf("packed decoder: na %u nb %u off %u\n", na, nb, off);
                    return 1;
                }
            }
        }
        for (uint32_t n = 0; n <= 18; n++) {  // the range masks: exactly the bits of the slots below n
            uint32_t want = 0;
            for (uint32_t i = 0; i < n; i++) want |= 1u << (10u * (i / 6) + i % 6);
            if ((cb_packed_below(n) & 0x03F0FC3Fu) != want) { std::printf("cb_packed_below(%u)\n", n); return 1; }
        }
    }
    std::printf("OK %llu cases, %llu ranks answered from records, %llu overflowed records\n", (unsigned long long)cases, (unsigned long long)answered, (unsigned long long)overflowed);
    return 0;
}
