// si_ref_wrap.cpp -- TEST INFRASTRUCTURE.  Thin extern "C" wrapper that lets tests/bench call the
// reference's own vendored, reference-native interval code: superintervals.hpp
// (/root/reference/datafusion/bio-function-ranges/superintervals/src/superintervals.hpp, selected by
// `SET bio.interval_join_algorithm = SuperIntervals`, IJ:832-845, IJ:892-898; asserted to give the same
// join rows as Coitrees in tests/integration_test.rs:93,189).  The header is NOT copied into this repo:
// oracle/Makefile compiles this file with -I pointing at the reference tree and writes only
// oracle/_ref/libsi_ref.so.  Used to cross-check the oracle's pair/count arithmetic and as a
// genuine-reference CPU baseline for overlap pairs.
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "superintervals.hpp"

struct si_ref_index {
    uint32_t n_keys;
    std::vector<si::IntervalMap<int32_t, uint32_t>> maps;
};

extern "C" {

si_ref_index* si_ref_build(const uint32_t* key, const int32_t* start, const int32_t* end, uint64_t n,
                           uint32_t n_keys) {
    auto* ix = new si_ref_index();
    ix->n_keys = n_keys;
    ix->maps.resize(n_keys);
    for (uint64_t i = 0; i < n; i++) {
        uint32_t c = key ? key[i] : 0;
        if (c < n_keys) ix->maps[c].add(start[i], end[i], (uint32_t)i);  // IJ:837-839
    }
    for (auto& m : ix->maps) m.build();  // IJ:840
    return ix;
}

void si_ref_destroy(si_ref_index* ix) { delete ix; }

// per-probe match counts via IntervalMap::count
void si_ref_counts(const si_ref_index* ix, const uint32_t* key, const int32_t* start, const int32_t* end,
                   uint64_t n, uint64_t* out, int threads) {
#pragma omp parallel for schedule(static, 8192) num_threads(threads > 0 ? threads : 1)
    for (uint64_t i = 0; i < n; i++) {
        uint32_t c = key ? key[i] : 0;
        if (c >= ix->n_keys) { out[i] = 0; continue; }
        auto& m = const_cast<si::IntervalMap<int32_t, uint32_t>&>(ix->maps[c]);
        out[i] = m.size() > 0 ? m.count(start[i], end[i]) : 0;
    }
}

// all (build_row, probe_row) pairs, probe-major (IJ:892-898 + IJ:1616-1653); arrays malloc'd
uint64_t si_ref_pairs(const si_ref_index* ix, const uint32_t* key, const int32_t* start, const int32_t* end,
                      uint64_t n, uint32_t** build_idx, uint32_t** probe_idx) {
    std::vector<uint32_t> b, p, found;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t c = key ? key[i] : 0;
        if (c >= ix->n_keys) continue;
        found.clear();
        ix->maps[c].search_values(start[i], end[i], found);
        for (uint32_t v : found) { b.push_back(v); p.push_back((uint32_t)i); }
    }
    *build_idx = (uint32_t*)malloc((b.size() ? b.size() : 1) * 4);
    *probe_idx = (uint32_t*)malloc((p.size() ? p.size() : 1) * 4);
    for (size_t j = 0; j < b.size(); j++) { (*build_idx)[j] = b[j]; (*probe_idx)[j] = p[j]; }
    return b.size();
}

// multi-threaded pair enumeration that only counts (CPU baseline timing without a giant output vector)
uint64_t si_ref_pairs_count_mt(const si_ref_index* ix, const uint32_t* key, const int32_t* start,
                               const int32_t* end, uint64_t n, int threads) {
    uint64_t total = 0;
#pragma omp parallel num_threads(threads > 0 ? threads : 1) reduction(+ : total)
    {
        std::vector<uint32_t> found;
#pragma omp for schedule(static, 8192)
        for (uint64_t i = 0; i < n; i++) {
            uint32_t c = key ? key[i] : 0;
            if (c >= ix->n_keys) continue;
            found.clear();
            ix->maps[c].search_values(start[i], end[i], found);
            total += found.size();
        }
    }
    return total;
}

void si_ref_free(void* p) { free(p); }
}
