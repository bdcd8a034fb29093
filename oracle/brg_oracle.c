/*
 * brg_oracle.c -- CPU restatement of the reference's interval-overlap hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (datafusion-bio-functions_b200/csrc) never links, loads or calls anything in oracle/.
 *
 * Every function restates the reference's arithmetic (paths relative to
 * /root/reference/datafusion/bio-function-ranges/src):
 *   IT = interval_tree.rs, NI = nearest_index.rs, NE = nearest.rs,
 *   IJ = physical_planner/joins/interval_join.rs
 * The third-party dependency coitrees 0.4.0 (Cargo.lock:824, not vendored) is restated
 * from its published behaviour: end-inclusive intervals, query(first,last) visits every
 * stored interval with node.first <= last && node.last >= first, in unspecified order,
 * with no guard for first > last.  Here it is an implicit max-end-augmented balanced
 * tree over the start-sorted array (a textbook structure, deliberately NOT the
 * algorithm the CUDA path uses, so the two check each other).
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against the reference's
 * unit-test literals (IT:355-412, NI:355-417), the CSV expected tables of
 * tests/integration_test.rs, the 438 694-row expected_coverage.parquet golden file,
 * and the reference-native superintervals.hpp compiled into oracle/_ref.
 * Unpinned (declared): degenerate probes (start > end after the strict adjustment) on
 * the coitrees-backed coverage/pairs paths -- no reference test covers them.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BRGO_ABSENT 0xFFFFFFFFu

typedef struct {
    int32_t first, last;
    uint32_t pos; /* row id in the concatenated build batch (IJ:920, NE:536) */
} rec_t;

/* per-key (contig) index: everything the reference keeps per AHashMap entry */
typedef struct {
    uint64_t n;
    /* CountOverlapIndex (IT:20-38): two independently sorted arrays */
    int32_t *starts, *ends;
    /* NearestIntervalIndex (NI:45-85) */
    rec_t *by_start;          /* sorted (first,last,pos) */
    int32_t *prefix_max_end;  /* NI:58-63 */
    rec_t *by_end;            /* sorted (last,first,pos) NI:77-82 */
    int32_t *tree_max;        /* implicit tree over by_start: max `last` of the subtree range */
    /* coverage: merged intervals (IT:52-73) + tree over them */
    uint64_t n_merged;
    rec_t *merged;
    int32_t *merged_tree_max;
} key_index_t;

typedef struct brgo_index {
    uint32_t n_keys;
    key_index_t *k;
    /* raw build columns, for distances (NE:288-296) */
    uint64_t n_rows;
    int32_t *raw_start, *raw_end;
} brgo_index;

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}
/* NI:50-55 */
static int cmp_by_start(const void *a, const void *b) {
    const rec_t *x = a, *y = b;
    if (x->first != y->first) return x->first < y->first ? -1 : 1;
    if (x->last != y->last) return x->last < y->last ? -1 : 1;
    return (x->pos > y->pos) - (x->pos < y->pos);
}
/* NI:77-82 */
static int cmp_by_end(const void *a, const void *b) {
    const rec_t *x = a, *y = b;
    if (x->last != y->last) return x->last < y->last ? -1 : 1;
    if (x->first != y->first) return x->first < y->first ? -1 : 1;
    return (x->pos > y->pos) - (x->pos < y->pos);
}

/* subtree over sorted slice [lo,hi): root = mid; returns max last; stores at tmax[mid] */
static int32_t build_tree(const rec_t *a, int32_t *tmax, int64_t lo, int64_t hi) {
    if (lo >= hi) return INT32_MIN;
    int64_t mid = lo + (hi - lo) / 2;
    int32_t m = a[mid].last;
    int32_t l = build_tree(a, tmax, lo, mid);
    int32_t r = build_tree(a, tmax, mid + 1, hi);
    if (l > m) m = l;
    if (r > m) m = r;
    tmax[mid] = m;
    return m;
}

typedef void (*visit_fn)(const rec_t *, void *);

/* coitrees-style query: visit all with first <= qlast && last >= qfirst (no guard) */
static void tree_query(const rec_t *a, const int32_t *tmax, int64_t lo, int64_t hi,
                       int32_t qfirst, int32_t qlast, visit_fn f, void *ud) {
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (tmax[mid] < qfirst) return; /* nothing in this subtree reaches the query */
        tree_query(a, tmax, lo, mid, qfirst, qlast, f, ud);
        if (a[mid].first > qlast) return; /* node and everything right of it start after */
        if (a[mid].last >= qfirst) f(&a[mid], ud);
        lo = mid + 1;
    }
}

/* IT:52-73 merge_intervals: stable sort by first; merge when next.first <= cur.last */
static uint64_t merge_intervals(const rec_t *sorted_by_first, uint64_t n, rec_t *out) {
    if (n == 0) return 0;
    uint64_t m = 0;
    rec_t cur = sorted_by_first[0];
    for (uint64_t i = 1; i < n; i++) {
        rec_t iv = sorted_by_first[i];
        if (iv.first <= cur.last) {
            if (iv.last > cur.last) cur.last = iv.last;
        } else {
            out[m++] = cur;
            cur = iv;
        }
    }
    out[m++] = cur;
    return m;
}

void brgo_index_destroy(brgo_index *ix) {
    if (!ix) return;
    for (uint32_t c = 0; c < ix->n_keys; c++) {
        key_index_t *k = &ix->k[c];
        free(k->starts); free(k->ends); free(k->by_start); free(k->prefix_max_end);
        free(k->by_end); free(k->tree_max); free(k->merged); free(k->merged_tree_max);
    }
    free(ix->k); free(ix->raw_start); free(ix->raw_end); free(ix);
}

/* Build: group rows by key id (the host-side dictionary id of the contig / `on` tuple:
 * IT:90-98, NE:511-530, IJ:903-931), then build every per-key structure. */
brgo_index *brgo_build(const uint32_t *key, const int32_t *start, const int32_t *end,
                       uint64_t n, uint32_t n_keys) {
    brgo_index *ix = calloc(1, sizeof(*ix));
    ix->n_keys = n_keys;
    ix->k = calloc(n_keys ? n_keys : 1, sizeof(key_index_t));
    ix->n_rows = n;
    ix->raw_start = malloc((n ? n : 1) * sizeof(int32_t));
    ix->raw_end = malloc((n ? n : 1) * sizeof(int32_t));
    if (n) { memcpy(ix->raw_start, start, n * 4); memcpy(ix->raw_end, end, n * 4); }
    uint64_t *cnt = calloc(n_keys + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) {
        uint32_t c = key ? key[i] : 0;
        if (c < n_keys) cnt[c]++;
    }
    for (uint32_t c = 0; c < n_keys; c++) {
        key_index_t *k = &ix->k[c];
        uint64_t m = cnt[c];
        k->n = 0;
        k->by_start = malloc((m ? m : 1) * sizeof(rec_t));
    }
    for (uint64_t i = 0; i < n; i++) {
        uint32_t c = key ? key[i] : 0;
        if (c >= n_keys) continue;
        key_index_t *k = &ix->k[c];
        rec_t r = {start[i], end[i], (uint32_t)i};
        k->by_start[k->n++] = r;
    }
    free(cnt);
#pragma omp parallel for schedule(dynamic, 1)
    for (uint32_t c = 0; c < n_keys; c++) {
        key_index_t *k = &ix->k[c];
        uint64_t m = k->n, a = m ? m : 1;
        k->starts = malloc(a * 4); k->ends = malloc(a * 4);
        for (uint64_t i = 0; i < m; i++) { k->starts[i] = k->by_start[i].first; k->ends[i] = k->by_start[i].last; }
        qsort(k->starts, m, 4, cmp_i32); /* IT:35 */
        qsort(k->ends, m, 4, cmp_i32);   /* IT:36 */
        k->by_end = malloc(a * sizeof(rec_t));
        memcpy(k->by_end, k->by_start, m * sizeof(rec_t));
        qsort(k->by_start, m, sizeof(rec_t), cmp_by_start);
        qsort(k->by_end, m, sizeof(rec_t), cmp_by_end);
        k->prefix_max_end = malloc(a * 4);
        int32_t mx = INT32_MIN;
        for (uint64_t i = 0; i < m; i++) { if (k->by_start[i].last > mx) mx = k->by_start[i].last; k->prefix_max_end[i] = mx; }
        k->tree_max = malloc(a * 4);
        build_tree(k->by_start, k->tree_max, 0, (int64_t)m);
        k->merged = malloc(a * sizeof(rec_t));
        /* by_start is sorted by first (ties by last,pos): merging is insensitive to tie order */
        k->n_merged = merge_intervals(k->by_start, m, k->merged);
        k->merged_tree_max = malloc(a * 4);
        build_tree(k->merged, k->merged_tree_max, 0, (int64_t)k->n_merged);
    }
    return ix;
}

/* slice::partition_point helpers */
static uint64_t pp_i32_le(const int32_t *a, uint64_t n, int32_t v) { /* #(a <= v) */
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = lo + (hi - lo) / 2; if (a[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
}
static uint64_t pp_i32_lt(const int32_t *a, uint64_t n, int32_t v) { /* #(a < v) */
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = lo + (hi - lo) / 2; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
static uint64_t pp_first_le(const rec_t *a, uint64_t n, int32_t v) { /* #(a.first <= v) */
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = lo + (hi - lo) / 2; if (a[mid].first <= v) lo = mid + 1; else hi = mid; }
    return lo;
}
static uint64_t pp_last_lt(const rec_t *a, uint64_t n, int32_t v) { /* #(a.last < v) */
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = lo + (hi - lo) / 2; if (a[mid].last < v) lo = mid + 1; else hi = mid; }
    return lo;
}

/* wrapping i32 arithmetic (Rust release builds wrap; IT:185-188, IT:146-151) */
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

static const key_index_t *lookup(const brgo_index *ix, const uint32_t *key, uint64_t i) {
    uint32_t c = key ? key[i] : 0;
    if (c >= ix->n_keys) return NULL; /* unknown contig: IT:265, NE:354 */
    return &ix->k[c];
}

/* IT:41-49 query_count + IT:249-267 per-row loop */
void brgo_count_overlaps(const brgo_index *ix, const uint32_t *key, const int32_t *start,
                         const int32_t *end, uint64_t n, int strict, int64_t *out, int threads) {
#pragma omp parallel for schedule(static, 8192) num_threads(threads > 0 ? threads : 1)
    for (uint64_t i = 0; i < n; i++) {
        int32_t qs = start[i], qe = end[i];
        if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); } /* IT:253-256 */
        const key_index_t *k = lookup(ix, key, i);
        int64_t c = 0;
        if (k && !(qe < qs)) { /* IT:42-44 */
            uint64_t started = pp_i32_le(k->starts, k->n, qe);      /* IT:46 */
            uint64_t ended_before = pp_i32_lt(k->ends, k->n, qs);   /* IT:47 */
            c = (int64_t)(started - ended_before);                  /* IT:48 (usize wrapping sub) */
        }
        out[i] = c;
    }
}

typedef struct { int32_t qs, qe, acc; } cov_ud;
/* IT:145-152 */
static void cov_visit(const rec_t *node, void *p) {
    cov_ud *u = p;
    int32_t hi = wadd(u->qe, 1) < node->last ? wadd(u->qe, 1) : node->last;   /* min(end+1,last) */
    int32_t lo = wsub(u->qs, 1) > node->first ? wsub(u->qs, 1) : node->first; /* max(start-1,first) */
    int32_t ov = wsub(hi, lo);
    if (ov < 1) ov = 1;
    u->acc = wadd(u->acc, ov);
}

/* IT:181-208 (coverage=true) */
void brgo_coverage(const brgo_index *ix, const uint32_t *key, const int32_t *start,
                   const int32_t *end, uint64_t n, int strict, int64_t *out, int threads) {
#pragma omp parallel for schedule(static, 8192) num_threads(threads > 0 ? threads : 1)
    for (uint64_t i = 0; i < n; i++) {
        int32_t qs = start[i], qe = end[i];
        if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); }
        const key_index_t *k = lookup(ix, key, i);
        cov_ud u = {qs, qe, 0};
        if (k) tree_query(k->merged, k->merged_tree_max, 0, (int64_t)k->n_merged, qs, qe, cov_visit, &u);
        out[i] = (int64_t)u.acc; /* IT:207 */
    }
}

/* NI:252-260 */
static int64_t candidate_distance(int32_t qs, int32_t qe, int32_t is, int32_t ie) {
    if (qe < is) return (int64_t)is - (int64_t)qe;
    if (ie < qs) return (int64_t)qs - (int64_t)ie;
    return 0;
}
/* NI:244-250 + NI:262-266 */
static int cmp_candidate(int32_t qs, int32_t qe, const rec_t *a, const rec_t *b) {
    int64_t ad = candidate_distance(qs, qe, a->first, a->last);
    int64_t bd = candidate_distance(qs, qe, b->first, b->last);
    if (ad != bd) return ad < bd ? -1 : 1;
    return cmp_by_start(a, b);
}
/* NI:222-234 */
static const rec_t *first_overlap_by_start(const key_index_t *k, int32_t qs, int32_t qe) {
    if (qe < qs) return NULL;
    uint64_t prefix_len = pp_first_le(k->by_start, k->n, qe);
    if (prefix_len == 0 || k->prefix_max_end[prefix_len - 1] < qs) return NULL;
    uint64_t idx = pp_i32_lt(k->prefix_max_end, prefix_len, qs);
    return &k->by_start[idx];
}
/* NI:192-220 */
static const rec_t *nearest_non_overlap_one(const key_index_t *k, int32_t qs, int32_t qe) {
    uint64_t left_idx = pp_last_lt(k->by_end, k->n, qs);
    uint64_t right_idx = pp_first_le(k->by_start, k->n, qe);
    const rec_t *left = left_idx > 0 ? &k->by_end[left_idx - 1] : NULL;
    const rec_t *right = right_idx < k->n ? &k->by_start[right_idx] : NULL;
    if (!left) return right;
    if (!right) return left;
    return cmp_candidate(qs, qe, left, right) <= 0 ? left : right;
}
/* NI:91-101 */
static const rec_t *nearest_one(const key_index_t *k, int32_t qs, int32_t qe, int include_overlaps) {
    if (k->n == 0) return NULL;
    if (include_overlaps) {
        const rec_t *b = first_overlap_by_start(k, qs, qe);
        if (b) return b;
    }
    return nearest_non_overlap_one(k, qs, qe);
}

/* NE:331-385 (k = 1).  left_idx[i] = build row (0 when invalid), valid[i] in {0,1},
 * dist[i] (may be NULL) from RAW coordinates NE:367-375; 0 when invalid (Arrow NULL). */
void brgo_nearest_k1(const brgo_index *ix, const uint32_t *key, const int32_t *start,
                     const int32_t *end, uint64_t n, int include_overlaps, int strict,
                     uint32_t *left_idx, uint8_t *valid, int64_t *dist, int threads) {
#pragma omp parallel for schedule(static, 8192) num_threads(threads > 0 ? threads : 1)
    for (uint64_t i = 0; i < n; i++) {
        int32_t qs = start[i], qe = end[i];
        if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); } /* NE:341-344 */
        const key_index_t *k = lookup(ix, key, i);
        const rec_t *b = k ? nearest_one(k, qs, qe, include_overlaps) : NULL;
        if (b) {
            left_idx[i] = b->pos; valid[i] = 1;
            if (dist) dist[i] = candidate_distance(start[i], end[i], ix->raw_start[b->pos], ix->raw_end[b->pos]);
        } else {
            left_idx[i] = 0; valid[i] = 0;
            if (dist) dist[i] = 0;
        }
    }
}

typedef struct { rec_t *v; uint64_t n, cap; } recvec;
static void recvec_push(recvec *r, rec_t x) {
    if (r->n == r->cap) { r->cap = r->cap ? r->cap * 2 : 16; r->v = realloc(r->v, r->cap * sizeof(rec_t)); }
    r->v[r->n++] = x;
}
static void collect_visit(const rec_t *node, void *p) { recvec_push((recvec *)p, *node); }

static int seen_insert(uint32_t *seen, uint64_t *n_seen, uint32_t pos) {
    for (uint64_t j = 0; j < *n_seen; j++) if (seen[j] == pos) return 0;
    seen[(*n_seen)++] = pos;
    return 1;
}

/* NI:103-190 nearest_k for one probe; writes up to k positions to out, returns count */
static uint64_t nearest_k_one(const key_index_t *kx, int32_t qs, int32_t qe, uint64_t k,
                              int include_overlaps, uint32_t *out, recvec *scratch) {
    if (k == 0 || kx->n == 0) return 0;
    if (k == 1) {
        const rec_t *b = nearest_one(kx, qs, qe, include_overlaps);
        if (b) { out[0] = b->pos; return 1; }
        return 0;
    }
    uint64_t n_out = 0; /* `seen` == the positions already in out (every insert is pushed) */
    if (include_overlaps) {
        scratch->n = 0;
        tree_query(kx->by_start, kx->tree_max, 0, (int64_t)kx->n, qs, qe, collect_visit, scratch);
        qsort(scratch->v, scratch->n, sizeof(rec_t), cmp_by_start); /* NI:129 */
        for (uint64_t j = 0; j < scratch->n; j++) {
            if (n_out == k) return n_out;
            seen_insert(out, &n_out, scratch->v[j].pos);
        }
    }
    if (n_out == k) return n_out;
    uint64_t left_idx = pp_last_lt(kx->by_end, kx->n, qs);
    uint64_t right_idx = pp_first_le(kx->by_start, kx->n, qe);
    while (n_out < k) {
        const rec_t *left = left_idx > 0 ? &kx->by_end[left_idx - 1] : NULL;
        const rec_t *right = right_idx < kx->n ? &kx->by_start[right_idx] : NULL;
        const rec_t *next;
        if (!left && !right) break;
        if (left && !right) { left_idx--; next = left; }
        else if (!left && right) { right_idx++; next = right; }
        else if (cmp_candidate(qs, qe, left, right) <= 0) { left_idx--; next = left; }
        else { right_idx++; next = right; }
        if (!include_overlaps && candidate_distance(qs, qe, next->first, next->last) == 0) continue; /* NI:182-184 */
        seen_insert(out, &n_out, next->pos);
    }
    return n_out;
}

/* NE:386-456 (k > 1, also valid for k = 1).  Outputs sized n*max(k,1):
 * left_idx, valid, right_idx (probe row), dist (nullable ptr).  Returns rows written. */
uint64_t brgo_nearest_k(const brgo_index *ix, const uint32_t *key, const int32_t *start,
                        const int32_t *end, uint64_t n, uint32_t k, int include_overlaps, int strict,
                        uint32_t *left_idx, uint8_t *valid, uint32_t *right_idx, int64_t *dist) {
    uint64_t w = 0;
    uint32_t *buf = malloc((k ? k : 1) * sizeof(uint32_t));
    recvec scratch = {0};
    for (uint64_t i = 0; i < n; i++) {
        int32_t qs = start[i], qe = end[i];
        if (strict) { qs = wadd(qs, 1); qe = wsub(qe, 1); }
        const key_index_t *kx = lookup(ix, key, i);
        uint64_t m = kx ? nearest_k_one(kx, qs, qe, k, include_overlaps, buf, &scratch) : 0;
        if (m == 0) { /* NE:433-439 */
            left_idx[w] = 0; valid[w] = 0; right_idx[w] = (uint32_t)i; if (dist) dist[w] = 0; w++;
        } else {
            for (uint64_t j = 0; j < m; j++) {
                left_idx[w] = buf[j]; valid[w] = 1; right_idx[w] = (uint32_t)i;
                if (dist) dist[w] = candidate_distance(start[i], end[i], ix->raw_start[buf[j]], ix->raw_end[buf[j]]);
                w++;
            }
        }
    }
    free(buf); free(scratch.v);
    return w;
}

typedef struct { uint32_t *v; uint64_t n, cap; } u32vec;
static void u32vec_push(u32vec *r, uint32_t x) {
    if (r->n == r->cap) { r->cap = r->cap ? r->cap * 2 : 1024; r->v = realloc(r->v, r->cap * sizeof(uint32_t)); }
    r->v[r->n++] = x;
}
static void pos_visit(const rec_t *node, void *p) { u32vec_push((u32vec *)p, node->pos); }
static void count_visit(const rec_t *node, void *p) { (void)node; (*(uint64_t *)p)++; }

/* IJ:854-863 + IJ:1616-1653: per probe row, every build position with
 * b.start <= q.end && b.end >= q.start (closed; start/end are whatever the host evaluated,
 * i.e. `end - 1` already applied for strict comparisons, PI:63-124).  Probe-major order.
 * Returns pair count; *build_idx / *probe_idx are malloc'd (free with brgo_free). */
uint64_t brgo_overlap_pairs(const brgo_index *ix, const uint32_t *key, const int32_t *start,
                            const int32_t *end, uint64_t n, uint32_t **build_idx, uint32_t **probe_idx) {
    u32vec b = {0}, p = {0};
    for (uint64_t i = 0; i < n; i++) {
        const key_index_t *k = lookup(ix, key, i);
        if (!k) continue;
        uint64_t before = b.n;
        tree_query(k->by_start, k->tree_max, 0, (int64_t)k->n, start[i], end[i], pos_visit, &b);
        for (uint64_t j = before; j < b.n; j++) u32vec_push(&p, (uint32_t)i);
    }
    *build_idx = b.v; *probe_idx = p.v;
    return b.n;
}

/* per-probe number of matches under the coitrees predicate (RightSemi/RightAnti: IJ:1433-1463) */
void brgo_overlap_counts(const brgo_index *ix, const uint32_t *key, const int32_t *start,
                         const int32_t *end, uint64_t n, uint64_t *out, int threads) {
#pragma omp parallel for schedule(static, 8192) num_threads(threads > 0 ? threads : 1)
    for (uint64_t i = 0; i < n; i++) {
        const key_index_t *k = lookup(ix, key, i);
        uint64_t c = 0;
        if (k) tree_query(k->by_start, k->tree_max, 0, (int64_t)k->n, start[i], end[i], count_visit, &c);
        out[i] = c;
    }
}

/* Order-independent checksum of the pair set (the scheme of the reference's large-table test, TEST:4289-4404: row hashes
 * combined by wrapping sum, wrapping sum of squares and XOR, so that any emission order gives the same triple).
 * h(build_row, probe_row) is brgo_pair_hash below; probe rows are numbered row_base + i.  out_counts may be NULL.
 * sums[0] = sum h, sums[1] = sum h*h (both mod 2^64), sums[2] = XOR h, sums[3] = number of pairs. */
static inline uint64_t brgo_pair_hash(uint32_t b, uint64_t p) {
    uint64_t h = (uint64_t)b * 0x9E3779B97F4A7C15ull + p * 0xC2B2AE3D27D4EB4Full + 0x165667B19E3779F9ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    return h;
}
typedef struct { uint64_t s1, s2, x, n, p; } csum_t;
static void csum_visit(const rec_t *node, void *ud) {
    csum_t *c = ud;
    const uint64_t h = brgo_pair_hash(node->pos, c->p);
    c->s1 += h; c->s2 += h * h; c->x ^= h; c->n++;
}
void brgo_overlap_checksum(const brgo_index *ix, const uint32_t *key, const int32_t *start, const int32_t *end, uint64_t n,
                           uint64_t row_base, uint64_t *out_counts, uint64_t *sums, int threads) {
    uint64_t S1 = 0, S2 = 0, X = 0, N = 0;
#pragma omp parallel for schedule(static, 8192) num_threads(threads > 0 ? threads : 1) reduction(+ : S1, S2, N) reduction(^ : X)
    for (uint64_t i = 0; i < n; i++) {
        const key_index_t *k = lookup(ix, key, i);
        csum_t c = {0, 0, 0, 0, row_base + i};
        if (k) tree_query(k->by_start, k->tree_max, 0, (int64_t)k->n, start[i], end[i], csum_visit, &c);
        if (out_counts) out_counts[i] = c.n;
        S1 += c.s1; S2 += c.s2; X ^= c.x; N += c.n;
    }
    sums[0] = S1; sums[1] = S2; sums[2] = X; sums[3] = N;
}

/* The pair join as the reference runs it on `threads` partitions (IJ:1616-1653 per probe batch of a partition): every thread
 * takes one contiguous range of the probe rows and MATERIALISES its own build-index vector and the run-length-expanded
 * probe-index vector, exactly like brgo_overlap_pairs above; nothing is concatenated (the reference hands each partition's
 * batches downstream as they are).  This is the CPU baseline of bench.py's pairs workload -- the work the reference does,
 * output vectors included -- and returns the pair count; sums (may be NULL) receives the order-independent checksum of the
 * materialised vectors in the layout of brgo_overlap_checksum, which is how tests/ know the vectors were right. */
uint64_t brgo_overlap_pairs_partitions(const brgo_index *ix, const uint32_t *key, const int32_t *start, const int32_t *end,
                                       uint64_t n, uint64_t row_base, int threads, uint64_t *sums) {
    const int T = threads > 0 ? threads : 1;
    uint64_t S1 = 0, S2 = 0, X = 0, N = 0;
#pragma omp parallel for schedule(static, 1) num_threads(T) reduction(+ : S1, S2, N) reduction(^ : X)
    for (int t = 0; t < T; t++) {
        const uint64_t lo = n * (uint64_t)t / (uint64_t)T, hi = n * (uint64_t)(t + 1) / (uint64_t)T;
        u32vec b = {0}, p = {0};
        for (uint64_t i = lo; i < hi; i++) {
            const key_index_t *k = lookup(ix, key, i);
            if (!k) continue;
            const uint64_t before = b.n;
            tree_query(k->by_start, k->tree_max, 0, (int64_t)k->n, start[i], end[i], pos_visit, &b);
            for (uint64_t j = before; j < b.n; j++) u32vec_push(&p, (uint32_t)i);
        }
        if (sums)
            for (uint64_t j = 0; j < b.n; j++) {
                const uint64_t h = brgo_pair_hash(b.v[j], row_base + p.v[j]);
                S1 += h; S2 += h * h; X ^= h;
            }
        N += b.n;
        free(b.v); free(p.v);
    }
    if (sums) { sums[0] = S1; sums[1] = S2; sums[2] = X; sums[3] = N; }
    return N;
}


/* IT:52-73 exposed for the unit-test literals (IT:291-353): input need not be sorted */
uint64_t brgo_merge_intervals(const int32_t *first, const int32_t *last, uint64_t n,
                              int32_t *out_first, int32_t *out_last) {
    rec_t *a = malloc((n ? n : 1) * sizeof(rec_t)), *m = malloc((n ? n : 1) * sizeof(rec_t));
    for (uint64_t i = 0; i < n; i++) { a[i].first = first[i]; a[i].last = last[i]; a[i].pos = (uint32_t)i; }
    /* Rust sort_by is stable on `first` only; tie order does not change the merged result */
    qsort(a, n, sizeof(rec_t), cmp_by_start);
    uint64_t k = merge_intervals(a, n, m);
    for (uint64_t i = 0; i < k; i++) { out_first[i] = m[i].first; out_last[i] = m[i].last; }
    free(a); free(m);
    return k;
}

/* ---- merge (merge.rs:260-358 + grouped_stream.rs:48-98): per contig, intervals sorted by (start, end); an interval joins
 * the current run when start <= cur_end + min_dist (strict: <), the boundary computed with a saturating i64 add;
 * output rows (contig, cur_start, cur_end, n_intervals), contigs in id order (the caller assigns ids in name order). */
typedef struct { int64_t s, e; } iv64_t;
static int cmp_iv64(const void *a, const void *b) {
    const iv64_t *x = a, *y = b;
    if (x->s != y->s) return x->s < y->s ? -1 : 1;
    if (x->e != y->e) return x->e < y->e ? -1 : 1;
    return 0;
}
static int64_t sat_add_i64(int64_t a, int64_t b) {
    if (b > 0 && a > INT64_MAX - b) return INT64_MAX;
    if (b < 0 && a < INT64_MIN - b) return INT64_MIN;
    return a + b;
}
uint64_t brgo_merge(const uint32_t *key, const int64_t *start, const int64_t *end, uint64_t n, uint32_t n_keys,
                    int64_t min_dist, int strict, uint32_t *out_key, int64_t *out_start, int64_t *out_end, int64_t *out_count) {
    uint64_t *cnt = calloc((size_t)n_keys + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) cnt[(key ? key[i] : 0) + 1]++;
    for (uint32_t k = 0; k < n_keys; k++) cnt[k + 1] += cnt[k];
    iv64_t *rows = malloc((n ? n : 1) * sizeof(iv64_t));
    uint64_t *fill = malloc(((size_t)n_keys + 1) * sizeof(uint64_t));
    memcpy(fill, cnt, ((size_t)n_keys + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) { iv64_t r = {start[i], end[i]}; rows[fill[key ? key[i] : 0]++] = r; }
    uint64_t m = 0;
    for (uint32_t k = 0; k < n_keys; k++) {
        iv64_t *g = rows + cnt[k];
        const uint64_t gn = cnt[k + 1] - cnt[k];
        if (!gn) continue;
        qsort(g, gn, sizeof(iv64_t), cmp_iv64);
        int64_t cs = g[0].s, ce = g[0].e, cc = 1;
        for (uint64_t i = 1; i < gn; i++) {
            const int64_t boundary = sat_add_i64(ce, min_dist);
            const int join = strict ? g[i].s < boundary : g[i].s <= boundary;
            if (join) { if (g[i].e > ce) ce = g[i].e; cc++; }
            else { out_key[m] = k; out_start[m] = cs; out_end[m] = ce; out_count[m] = cc; m++; cs = g[i].s; ce = g[i].e; cc = 1; }
        }
        out_key[m] = k; out_start[m] = cs; out_end[m] = ce; out_count[m] = cc; m++;
    }
    free(cnt); free(rows); free(fill);
    return m;
}

/* ---- cluster (cluster.rs:560-720; ids: ClusterIdCoordinator::build_offsets, cluster.rs:383-414): same sweep as merge, one
 * output row per input row in (contig id, start, end, row) order; cluster ids are numbered globally in contig order. */
typedef struct { int64_t s, e; uint64_t row; } ivr64_t;
static int cmp_ivr64(const void *a, const void *b) {
    const ivr64_t *x = a, *y = b;
    if (x->s != y->s) return x->s < y->s ? -1 : 1;
    if (x->e != y->e) return x->e < y->e ? -1 : 1;
    return x->row < y->row ? -1 : (x->row > y->row ? 1 : 0);
}
void brgo_cluster(const uint32_t *key, const int64_t *start, const int64_t *end, uint64_t n, uint32_t n_keys, int64_t min_dist,
                  int strict, uint64_t *out_row, int64_t *out_cluster, int64_t *out_cstart, int64_t *out_cend) {
    uint64_t *cnt = calloc((size_t)n_keys + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) cnt[(key ? key[i] : 0) + 1]++;
    for (uint32_t k = 0; k < n_keys; k++) cnt[k + 1] += cnt[k];
    ivr64_t *rows = malloc((n ? n : 1) * sizeof(ivr64_t));
    uint64_t *fill = malloc(((size_t)n_keys + 1) * sizeof(uint64_t));
    memcpy(fill, cnt, ((size_t)n_keys + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) { ivr64_t r = {start[i], end[i], i}; rows[fill[key ? key[i] : 0]++] = r; }
    uint64_t o = 0;
    int64_t cluster_id = 0;
    for (uint32_t k = 0; k < n_keys; k++) {
        ivr64_t *g = rows + cnt[k];
        const uint64_t gn = cnt[k + 1] - cnt[k];
        if (!gn) continue;
        qsort(g, gn, sizeof(ivr64_t), cmp_ivr64);
        uint64_t first = 0;  /* first row of the pending cluster */
        int64_t cs = g[0].s, ce = g[0].e;
        for (uint64_t i = 1; i <= gn; i++) {
            int join = 0;
            if (i < gn) {
                const int64_t boundary = sat_add_i64(ce, min_dist);
                join = strict ? g[i].s < boundary : g[i].s <= boundary;
            }
            if (join) { if (g[i].e > ce) ce = g[i].e; continue; }
            for (uint64_t j = first; j < i; j++) { out_row[o] = g[j].row; out_cluster[o] = cluster_id; out_cstart[o] = cs; out_cend[o] = ce; o++; }
            cluster_id++;
            if (i < gn) { first = i; cs = g[i].s; ce = g[i].e; }
        }
    }
    free(cnt); free(rows); free(fill);
}

/* ---- subtract (subtract.rs:395-440 / 560-640): per contig, left intervals in (start, end, row) order, right intervals in
 * (start, end) order, the reference's cursor sweep restated literally (persistent right_cursor included).
 * Returns the number of pieces; only the first `cap` are written (call once with cap = 0 to size the outputs). */
uint64_t brgo_subtract(const uint32_t *lkey, const int64_t *lstart, const int64_t *lend, uint64_t nl, const uint32_t *rkey,
                       const int64_t *rstart, const int64_t *rend, uint64_t nr, uint32_t n_keys, int strict,
                       uint64_t *out_row, int64_t *out_start, int64_t *out_end, uint64_t cap) {
    uint64_t *lc = calloc((size_t)n_keys + 2, sizeof(uint64_t)), *rc = calloc((size_t)n_keys + 2, sizeof(uint64_t));
    for (uint64_t i = 0; i < nl; i++) lc[lkey[i] + 1]++;
    for (uint64_t i = 0; i < nr; i++) if (rkey[i] < n_keys) rc[rkey[i] + 1]++;
    for (uint32_t k = 0; k < n_keys; k++) { lc[k + 1] += lc[k]; rc[k + 1] += rc[k]; }
    ivr64_t *L = malloc((nl ? nl : 1) * sizeof(ivr64_t));
    iv64_t *R = malloc((nr ? nr : 1) * sizeof(iv64_t));
    uint64_t *lf = malloc(((size_t)n_keys + 1) * sizeof(uint64_t)), *rf = malloc(((size_t)n_keys + 1) * sizeof(uint64_t));
    memcpy(lf, lc, ((size_t)n_keys + 1) * sizeof(uint64_t));
    memcpy(rf, rc, ((size_t)n_keys + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < nl; i++) { ivr64_t r = {lstart[i], lend[i], i}; L[lf[lkey[i]]++] = r; }
    for (uint64_t i = 0; i < nr; i++) if (rkey[i] < n_keys) { iv64_t r = {rstart[i], rend[i]}; R[rf[rkey[i]]++] = r; }
    uint64_t o = 0;
    for (uint32_t k = 0; k < n_keys; k++) {
        ivr64_t *lg = L + lc[k];
        iv64_t *rg = R + rc[k];
        const uint64_t ln = lc[k + 1] - lc[k], rn = rc[k + 1] - rc[k];
        qsort(lg, ln, sizeof(ivr64_t), cmp_ivr64);
        qsort(rg, rn, sizeof(iv64_t), cmp_iv64);
        uint64_t right_cursor = 0;
        for (uint64_t i = 0; i < ln; i++) {
            const int64_t ls = lg[i].s, le = lg[i].e;
            while (right_cursor < rn && (strict ? rg[right_cursor].e <= ls : rg[right_cursor].e < ls)) right_cursor++;
            int64_t cursor = ls;
            for (uint64_t j = right_cursor; j < rn; j++) {
                const int64_t rs = rg[j].s, re = rg[j].e;
                if (strict ? rs >= le : rs > le) break;
                if (rs > cursor) { if (o < cap) { out_row[o] = lg[i].row; out_start[o] = cursor; out_end[o] = rs; } o++; }
                if (re > cursor) cursor = re;
            }
            if (cursor < le) { if (o < cap) { out_row[o] = lg[i].row; out_start[o] = cursor; out_end[o] = le; } o++; }
        }
    }
    free(lc); free(rc); free(L); free(R); free(lf); free(rf);
    return o;
}

/* ---- complement (complement.rs:297-358, 396-457): per contig, the input intervals sorted by (start, end)
 * (grouped_stream.rs:92-94) and merged with ComplementStream::merge_intervals (s <= cur_end joins; strict: s < cur_end),
 * then emit_contig_complement restated literally for each view in the order given (the caller lists a contig's views in
 * (start, end) order, as StreamCollector leaves them).  A view whose key has no input rows is a "trailing" view-only contig
 * (complement.rs:437-457): emitted whole, unconditionally.  Returns the number of gaps; the first `cap` are written. */
uint64_t brgo_complement(const uint32_t *key, const int64_t *start, const int64_t *end, uint64_t n, uint32_t n_keys,
                         const uint32_t *vkey, const int64_t *vstart, const int64_t *vend, uint64_t nv, int strict,
                         uint64_t *out_view, int64_t *out_start, int64_t *out_end, uint64_t cap) {
    uint64_t *cnt = calloc((size_t)n_keys + 2, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) cnt[(key ? key[i] : 0) + 1]++;
    for (uint32_t k = 0; k < n_keys; k++) cnt[k + 1] += cnt[k];
    iv64_t *rows = malloc((n ? n : 1) * sizeof(iv64_t));
    uint64_t *fill = malloc(((size_t)n_keys + 1) * sizeof(uint64_t));
    uint64_t *mcnt = calloc((size_t)n_keys + 1, sizeof(uint64_t)); /* merged intervals of key k: rows[cnt[k] .. cnt[k] + mcnt[k]) */
    memcpy(fill, cnt, ((size_t)n_keys + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) { iv64_t r = {start[i], end[i]}; rows[fill[key ? key[i] : 0]++] = r; }
    for (uint32_t k = 0; k < n_keys; k++) {
        iv64_t *g = rows + cnt[k];
        const uint64_t gn = cnt[k + 1] - cnt[k];
        if (!gn) continue;
        qsort(g, gn, sizeof(iv64_t), cmp_iv64);
        uint64_t m = 0; /* merge in place (complement.rs:297-320) */
        int64_t cs = g[0].s, ce = g[0].e;
        for (uint64_t i = 1; i < gn; i++) {
            const int join = strict ? g[i].s < ce : g[i].s <= ce;
            if (join) { if (g[i].e > ce) ce = g[i].e; }
            else { g[m].s = cs; g[m].e = ce; m++; cs = g[i].s; ce = g[i].e; }
        }
        g[m].s = cs; g[m].e = ce; m++;
        mcnt[k] = m;
    }
    uint64_t o = 0;
    for (uint64_t v = 0; v < nv; v++) {
        const int64_t view_start = vstart[v], view_end = vend[v];
        const uint32_t k = vkey[v];
        if (k >= n_keys || cnt[k + 1] == cnt[k]) { /* view-only contig */
            if (o < cap) { out_view[o] = v; out_start[o] = view_start; out_end[o] = view_end; }
            o++;
            continue;
        }
        const iv64_t *mg = rows + cnt[k];
        int64_t cursor = view_start; /* complement.rs:330-356 */
        for (uint64_t j = 0; j < mcnt[k]; j++) {
            const int64_t ms = mg[j].s, me = mg[j].e;
            if (me <= view_start) continue;
            if (ms >= view_end) break;
            const int64_t interval_start = ms > view_start ? ms : view_start;
            const int64_t interval_end = me < view_end ? me : view_end;
            if (interval_start > cursor) { if (o < cap) { out_view[o] = v; out_start[o] = cursor; out_end[o] = interval_start; } o++; }
            cursor = interval_end;
        }
        if (cursor < view_end) { if (o < cap) { out_view[o] = v; out_start[o] = cursor; out_end[o] = view_end; } o++; }
    }
    free(cnt); free(rows); free(fill); free(mcnt);
    return o;
}

void brgo_free(void *p) { free(p); }
int brgo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
