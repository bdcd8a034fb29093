/*
 * bioranges.h -- C ABI of the B200-native interval-overlap engine (libbioranges.so).
 *
 * This is the drop-in boundary for the interval-overlap hot path of datafusion-bio-function-ranges.
 * The reference has no FFI seam today; each entry point below replaces the internal Rust seam cited
 * next to it (paths relative to /root/reference/datafusion/bio-function-ranges/src).  The Rust host
 * code (IntervalJoinExec / CountOverlapsExec / NearestExec) keeps its DataFusion trait surface and calls
 * these functions through `extern "C"` (binding shown in INTEGRATION.md).
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types.  All functions return 0 (BRG_OK) or a brg_status;
 *    the message of the last failure on the calling thread is brg_last_error().  Nothing throws or aborts.
 *  - coordinates are i32, closed [start,end] (the reference resolves every position column to &[i32],
 *    array_utils.rs:73-131).  `key_id` is the host-side dense dictionary id (0..n_keys-1) of the contig /
 *    join-key tuple; BRG_KEY_ABSENT on the probe side means "key not present in the build side"
 *    (HashMap::get -> None: interval_tree.rs:265, nearest.rs:354, interval_join.rs:857).  key_id may be
 *    NULL when there is a single key (all rows key 0).
 *  - "host" entry points take caller-owned host buffers (Arrow value buffers; pinned or pageable) and
 *    include the host<->device copies; "_dev" entry points take device pointers, enqueue on the given
 *    cudaStream_t (passed as void*) and do not synchronise.  Pageable host buffers (ordinary heap memory, what
 *    arrow-rs allocates) are detected and moved through a pinned staging ring by the context's host threads
 *    (BRG_HOST_THREADS, default min(8, cores)), so the PCIe copies stay asynchronous; buffers from brg_pinned_alloc
 *    or cudaHostRegister are used in place.  A per-row probe key column is scanned for runs chunk by chunk
 *    (the reference's contig-cache loop, interval_tree.rs:258-265): a coordinate-sorted table sends a few runs per
 *    8M-row chunk instead of 4 bytes per row.
 *  - results that the library owns (brg_rows, brg_merged, brg_gaps) may be released after brg_ctx_destroy of the
 *    context that produced them; an index must be destroyed before its context.
 *  - brg_index is immutable after build and may be probed concurrently from many threads
 *    (the reference shares it as Arc across partitions: joins/utils.rs:73-107).
 *  - There is no CPU fallback: every entry point fails with BRG_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef BIORANGES_H
#define BIORANGES_H

#include <stdint.h>

#if defined(__GNUC__)
#define BRG_API __attribute__((visibility("default")))
#else
#define BRG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BRG_KEY_ABSENT 0xFFFFFFFFu

typedef enum {
    BRG_OK = 0,
    BRG_ERR_INVALID = 1,     /* bad argument */
    BRG_ERR_CUDA = 2,        /* CUDA runtime failure / no device */
    BRG_ERR_OOM = 3,         /* host or device allocation failed */
    BRG_ERR_OVERFLOW = 4,    /* row count exceeds UInt32 index capacity (nearest.rs:357-361) */
    BRG_ERR_UNSUPPORTED = 5
} brg_status;

/* which structures brg_build materialises (bitmask) */
enum {
    BRG_WANT_COUNT = 1,    /* sorted starts / sorted ends            (interval_tree.rs:20-38)   */
    BRG_WANT_COVERAGE = 2, /* merged intervals + width prefix sums   (interval_tree.rs:52-111)  */
    BRG_WANT_NEAREST = 4,  /* by_start, prefix_max_end, by_end       (nearest_index.rs:45-85)   */
    BRG_WANT_PAIRS = 8,    /* start-sorted rows + prefix_max_end     (interval_join.rs:748-763) */
    BRG_WANT_ALL = 15
};

/* probe-side modes of brg_overlap_pairs (interval_join.rs:1014-1048, 1433-1463, 1614-1653) */
enum {
    BRG_PAIRS_INNER = 0, /* every (build_row, probe_row) pair, probe-major                      */
    BRG_PAIRS_SEMI = 1,  /* probe rows with >= 1 match (RightSemi); build_idx is not produced   */
    BRG_PAIRS_ANTI = 2,  /* probe rows with no match   (RightAnti); build_idx is not produced   */
    BRG_PAIRS_INNER_ANY_ORDER = 3 /* INNER through the one-launch join (brg_pairs_join_dev): the same pairs; the pairs of one
                                   * chunk of brg_pairs_join_chunk_rows() probe rows are contiguous and probe-major, the chunks
                                   * come in no particular order (the reference's row order is unspecified too: coitrees visit
                                   * order, batch arrival order across partitions)                        */
};
/* flags of brg_pairs_join_dev */
enum { BRG_JOIN_ORDERED = 1 /* probe-major output throughout (a look-back over the chunks' counts instead of one atomic per chunk) */ };

typedef struct brg_ctx brg_ctx;     /* one per GPU: streams, staging buffers, scratch */
typedef struct brg_index brg_index; /* immutable build-side index resident in HBM      */
typedef struct brg_rows brg_rows;   /* variable-length result in library-owned pinned host memory */

BRG_API const char* brg_last_error(void);
BRG_API int brg_version(void);

/* ---- context ------------------------------------------------------------------------------------ */
BRG_API int brg_ctx_create(int device_id, brg_ctx** out);
BRG_API void brg_ctx_destroy(brg_ctx* ctx);
BRG_API int brg_ctx_device(const brg_ctx* ctx);
BRG_API int brg_ctx_sm_count(const brg_ctx* ctx);

/* pinned host memory for callers that want zero-staging transfers (Arrow buffers can be allocated here) */
BRG_API int brg_pinned_alloc(uint64_t bytes, void** out);
BRG_API void brg_pinned_free(void* p);

/* ---- build (replaces collect_left_input + IntervalJoinAlgorithm::new, interval_join.rs:584-668,746-772;
 *             build_count_index_from_batches / build_coitree_from_batches, interval_tree.rs:75-143;
 *             build_nearest_indexes + NearestIntervalIndex::from_records, nearest.rs:498-547,
 *             nearest_index.rs:45-85).  Row i of the input is build row id i (`position`). -------------- */
BRG_API int brg_build(brg_ctx* ctx, const uint32_t* key_id, const int32_t* start, const int32_t* end, uint64_t n_rows,
              uint32_t n_keys, uint32_t want, brg_index** out);
BRG_API int brg_build_dev(brg_ctx* ctx, const uint32_t* d_key_id, const int32_t* d_start, const int32_t* d_end,
                  uint64_t n_rows, uint32_t n_keys, uint32_t want, void* stream, brg_index** out);
BRG_API void brg_index_destroy(brg_index* idx);
BRG_API uint64_t brg_index_rows(const brg_index* idx);        /* build rows with a valid key */
BRG_API uint64_t brg_index_merged_rows(const brg_index* idx); /* merged (coverage) intervals  */
BRG_API uint64_t brg_index_device_bytes(const brg_index* idx);
/* Diagnostics: the bins tables the index holds for UNSORTED probe tables (one 32-byte record per coordinate bin, read by the
 * one-row-per-thread kernels).  table: 0 = count bins over (start, end) order, 1 = coverage bins over the merged intervals,
 * 2 = nearest / pair bins over (start, prefix-max end).  Returns the record count (0 = table absent); *bin_shift = log2 of the
 * bin width, *format = 0 (12 slots of 16 bits) or 1 (18 slots of 10 bits).  Either pointer may be NULL.  No reference
 * counterpart: the reference's index is the coitrees / sorted-array structure of interval_tree.rs:75-143. */
BRG_API uint64_t brg_index_bins_info(const brg_index* idx, int table, uint32_t* bin_shift, uint32_t* format);

/* Replication across GPUs (one process per GPU): the index is one device slab plus a small POD header.
 * Rank 0 exports, every other rank imports the header, receives the slab with ncclBroadcast
 * (torch.distributed.broadcast on a tensor wrapping the slab) -- mirrors PartitionMode::CollectLeft,
 * interval_join.rs:396-399,466-480. */
BRG_API int brg_index_export(const brg_index* idx, void* header_buf, uint64_t header_cap, uint64_t* header_len,
                     void** d_slab, uint64_t* slab_bytes);
BRG_API int brg_index_import(brg_ctx* ctx, const void* header, uint64_t header_len, brg_index** out, void** d_slab,
                     uint64_t* slab_bytes);

/* The slab starts with a copy of its own header, so a replica needs nothing but the slab bytes.  One process per GPU:
 * broadcast brg_index_device_bytes (8 bytes), then every other rank does brg_index_slab_alloc, receives the slab with ONE
 * ncclBroadcast into it and calls brg_index_attach (which reads the header back from the front of the slab, after the work
 * queued on `stream`, and takes ownership of the buffer; on failure the caller frees it with brg_index_slab_free). */
BRG_API int brg_index_slab_alloc(brg_ctx* ctx, uint64_t slab_bytes, void** d_slab);
BRG_API void brg_index_slab_free(brg_ctx* ctx, void* d_slab);
BRG_API int brg_index_attach(brg_ctx* ctx, void* d_slab, uint64_t slab_bytes, void* stream, brg_index** out);
/* One process driving several GPUs (a DataFusion ExecutionPlan whose execute(partition) maps partitions to devices,
 * interval_join.rs:442-550, with the build side shared by all partitions, joins/utils.rs:73-107): out[i] receives a replica
 * of `src` on the device of dst[i] (one brg_ctx per GPU).  The slab is copied device to device -- NVLink peer copies where
 * the topology has them -- in a doubling tree.  Replicas are independent indexes: destroy each with brg_index_destroy.
 * dst[i] may be on the source's own device (a plain device copy). */
BRG_API int brg_index_replicate(const brg_index* src, brg_ctx* const* dst, int n_dst, brg_index** out);

/* ---- count_overlaps (replaces CountOverlapIndex::query_count + get_count_stream's row loop,
 *      interval_tree.rs:41-49, 224-280).  strict => probe start+1 / end-1 (interval_tree.rs:253-256). ---- */
BRG_API int brg_count_overlaps(brg_ctx* ctx, const brg_index* idx, const uint32_t* key_id, const int32_t* start,
                       const int32_t* end, uint64_t n, int strict, int64_t* out_count);
BRG_API int brg_count_overlaps_dev(brg_ctx* ctx, const brg_index* idx, const uint32_t* d_key_id, const int32_t* d_start,
                           const int32_t* d_end, uint64_t n, int strict, int64_t* d_out_count, void* stream);

/* Run-length form of the probe key column.  The reference's row loops resolve the contig of a row by comparing it with
 * the previous row's (cached_contig, interval_tree.rs:258-265; nearest.rs:346-352): on a coordinate-sorted table that
 * loop produces a handful of runs, not a column.  run r covers probe rows [run_end[r-1], run_end[r]) (run_end ascending,
 * run_end[n_runs-1] == n) with key run_key[r] (BRG_KEY_ABSENT allowed).  Only the runs cross PCIe; the per-row key
 * column the kernels read is expanded on the device.  Saves a third of the host->device bytes of a probe call. */
BRG_API int brg_count_overlaps_runs(brg_ctx* ctx, const brg_index* idx, const uint32_t* run_key, const uint64_t* run_end,
                            uint64_t n_runs, const int32_t* start, const int32_t* end, uint64_t n, int strict,
                            int64_t* out_count);
BRG_API int brg_coverage_runs(brg_ctx* ctx, const brg_index* idx, const uint32_t* run_key, const uint64_t* run_end,
                      uint64_t n_runs, const int32_t* start, const int32_t* end, uint64_t n, int strict,
                      int64_t* out_coverage);
BRG_API int brg_nearest_k1_runs(brg_ctx* ctx, const brg_index* idx, const uint32_t* run_key, const uint64_t* run_end,
                        uint64_t n_runs, const int32_t* start, const int32_t* end, uint64_t n, int include_overlaps,
                        int strict, uint32_t* out_left_idx, uint8_t* out_valid_bits, int64_t* out_distance);

/* ---- Arrow ingest (replaces the host copies of PosArray::resolve, array_utils.rs:73-131).  The start /
 *      end VALUE BUFFERS of the probe batch are passed in their native Arrow width and cross PCIe as they are; the checked
 *      narrowing to i32 (i32::try_from per row) runs on the device in front of the probe kernel.  A value that does not fit
 *      fails the call with BRG_ERR_OVERFLOW and the reference's text "coordinate value {v} at row {i} overflows i32 (max
 *      2147483647)" (start column first, array_utils.rs:94-98); when `bad_rows` is not NULL it receives the first offending
 *      row of the start column and of the end column (UINT64_MAX = none) so that a caller that coalesced several record
 *      batches can report the batch-local row.  Keys travel as runs (see brg_count_overlaps_runs). */
enum { BRG_POS_INT32 = 0, BRG_POS_INT64 = 1, BRG_POS_UINT32 = 2, BRG_POS_UINT64 = 3 };
BRG_API int brg_count_overlaps_arrow(brg_ctx* ctx, const brg_index* idx, const uint32_t* run_key, const uint64_t* run_end,
                             uint64_t n_runs, const void* start, int start_type, const void* end, int end_type, uint64_t n,
                             int strict, int64_t* out_count, uint64_t* bad_rows /* [2] or NULL */);
BRG_API int brg_coverage_arrow(brg_ctx* ctx, const brg_index* idx, const uint32_t* run_key, const uint64_t* run_end,
                       uint64_t n_runs, const void* start, int start_type, const void* end, int end_type, uint64_t n,
                       int strict, int64_t* out_coverage, uint64_t* bad_rows /* [2] or NULL */);
BRG_API int brg_nearest_k1_arrow(brg_ctx* ctx, const brg_index* idx, const uint32_t* run_key, const uint64_t* run_end,
                         uint64_t n_runs, const void* start, int start_type, const void* end, int end_type, uint64_t n,
                         int include_overlaps, int strict, uint32_t* out_left_idx, uint8_t* out_valid_bits,
                         int64_t* out_distance, uint64_t* bad_rows /* [2] or NULL */);

/* ---- coverage (replaces get_coverage + get_stream's row loop, interval_tree.rs:145-152, 155-221) ------ */
BRG_API int brg_coverage(brg_ctx* ctx, const brg_index* idx, const uint32_t* key_id, const int32_t* start,
                 const int32_t* end, uint64_t n, int strict, int64_t* out_coverage);
BRG_API int brg_coverage_dev(brg_ctx* ctx, const brg_index* idx, const uint32_t* d_key_id, const int32_t* d_start,
                     const int32_t* d_end, uint64_t n, int strict, int64_t* d_out_coverage, void* stream);

/* ---- nearest, k = 1 (replaces NearestIntervalIndex::nearest_one + the k==1 row loop, nearest_index.rs:91-101,
 *      192-234; nearest.rs:331-385; also the CoitreesNearest join arm, interval_join.rs:864-870, 1627-1636).
 *      One output row per probe row.  out_left_idx[i] = build row (0 when there is no match);
 *      out_valid_bits = Arrow validity bitmap (LSB-first, ceil(n/8) bytes, padded to 4 bytes) shared by the
 *      left columns and the distance column; out_distance (may be NULL) from RAW probe coordinates
 *      (nearest.rs:367-375), 0 where invalid. */
/* include_overlaps: 0 = neighbours only, 1 = first overlap else nearest neighbour (the reference's two settings),
 * BRG_NEAREST_OVERLAP_ONLY = first overlap in (start, end, row id) order else NULL -- the one-row-per-probe mode of the
 * CoitreesCountOverlaps join arm (interval_join.rs:1627-1636; the reference emits the first interval its tree VISITS, an
 * order it leaves unspecified, so which of several overlapping intervals is reported is not pinned). */
#define BRG_NEAREST_OVERLAP_ONLY 2
BRG_API int brg_nearest_k1(brg_ctx* ctx, const brg_index* idx, const uint32_t* key_id, const int32_t* start,
                   const int32_t* end, uint64_t n, int include_overlaps, int strict, uint32_t* out_left_idx,
                   uint8_t* out_valid_bits, int64_t* out_distance);
BRG_API int brg_nearest_k1_dev(brg_ctx* ctx, const brg_index* idx, const uint32_t* d_key_id, const int32_t* d_start,
                       const int32_t* d_end, uint64_t n, int include_overlaps, int strict, uint32_t* d_out_left_idx,
                       uint32_t* d_out_valid_bits, int64_t* d_out_distance, void* stream);

/* ---- nearest, any k >= 1 (replaces nearest_k + the k>1 row loop, nearest_index.rs:103-190, nearest.rs:386-456).
 *      Up to k rows per probe row, one NULL row when there is none.  Result columns: build_idx, valid bitmap,
 *      probe_idx, distance (when compute_distance). */
BRG_API int brg_nearest(brg_ctx* ctx, const brg_index* idx, const uint32_t* key_id, const int32_t* start, const int32_t* end,
                uint64_t n, uint32_t k, int include_overlaps, int compute_distance, int strict, brg_rows** out);

/* ---- overlap pairs (replaces IntervalJoinAlgorithm::get (Coitrees arm) + RLE expansion,
 *      interval_join.rs:849-863, 1614-1653; RightSemi/RightAnti 1433-1463).  start/end are the columns the
 *      host evaluated (for `<`/`>` predicates the host has already applied `end - 1`, intervals.rs:63-124).
 *      max_rows > 0 caps the output like the low-memory mode (interval_join.rs:1466-1573): the call consumes
 *      the longest prefix of probe rows whose pairs fit (at least one row) and reports it in
 *      *probe_rows_consumed; probe_idx values are relative to the start of this call's probe arrays. */
BRG_API int brg_overlap_pairs(brg_ctx* ctx, const brg_index* idx, const uint32_t* key_id, const int32_t* start,
                      const int32_t* end, uint64_t n, int mode, uint64_t max_rows, brg_rows** out,
                      uint64_t* probe_rows_consumed);
/* device two-pass form: count -> exclusive scan (d_offsets[n+1], u64) ; then emit into caller buffers */
BRG_API int brg_pairs_count_dev(brg_ctx* ctx, const brg_index* idx, const uint32_t* d_key_id, const int32_t* d_start,
                        const int32_t* d_end, uint64_t n, uint64_t* d_offsets, void* stream);
BRG_API int brg_pairs_emit_dev(brg_ctx* ctx, const brg_index* idx, const uint32_t* d_key_id, const int32_t* d_start,
                       const int32_t* d_end, uint64_t n, const uint64_t* d_offsets, uint32_t* d_build_idx,
                       uint32_t* d_probe_idx, void* stream);

/* chunked device form (the fast path): the unit of count -> scan -> emit is a chunk of brg_pairs_chunk_rows() (256)
 * probe rows.  Pass 1 writes d_chunk_base (ceil(n/256)+1 u64: [c] = first output slot of chunk c, [last] = total pairs)
 * and d_row_ranges (n u64: the candidate range of every probe row in the start-sorted build rows).  Pass 2 needs only
 * the probe start column and those two arrays: it flattens each chunk's candidates across a warp and writes the pairs
 * as contiguous runs (probe-major; matches of one probe in (start, end, row id) order). */
BRG_API uint32_t brg_pairs_chunk_rows(void);
BRG_API int brg_pairs_chunk_offsets_dev(brg_ctx* ctx, const brg_index* idx, const uint32_t* d_key_id, const int32_t* d_start,
                                const int32_t* d_end, uint64_t n, uint64_t* d_chunk_base, uint64_t* d_row_ranges,
                                void* stream);
BRG_API int brg_pairs_fill_dev(brg_ctx* ctx, const brg_index* idx, const int32_t* d_start, uint64_t n,
                       const uint64_t* d_chunk_base, const uint64_t* d_row_ranges, uint32_t* d_build_idx,
                       uint32_t* d_probe_idx, void* stream);

/* the join in ONE launch: per chunk of brg_pairs_join_chunk_rows() probe rows the candidate ranges, the pair count, the chunk's
 * output slots and the emission, with nothing but the pairs written to HBM (replaces IntervalJoinAlgorithm::get + the RLE
 * expansion, interval_join.rs:849-863, 1616-1653).  The output size is not known before the launch: the caller passes buffers
 * of `capacity` pairs; *d_total (device u64) receives the exact pair count, pairs beyond `capacity` are dropped (call again
 * with larger buffers when *d_total > capacity).  Output order: the pairs of a chunk are contiguous and probe-major (matches
 * of one probe in (start, end, row id) order); the chunks land in the order they finish unless flags has BRG_JOIN_ORDERED
 * (then the whole output is probe-major, as in the chunked form, at about twice the kernel time: every chunk has to wait
 * for the counts of all earlier ones).  d_chunk_slots (may be NULL): 2 u64 per chunk = {first output slot, pairs}. */
BRG_API uint32_t brg_pairs_join_chunk_rows(void);
BRG_API int brg_pairs_join_dev(brg_ctx* ctx, const brg_index* idx, const uint32_t* d_key_id, const int32_t* d_start,
                       const int32_t* d_end, uint64_t n, uint64_t capacity, int flags, uint32_t* d_build_idx,
                       uint32_t* d_probe_idx, uint64_t* d_total, uint64_t* d_chunk_slots, void* stream);

/* ---- output materialisation on the device (replaces arrow::compute::take of a fixed-width column by the emitted
 *      index column, interval_join.rs:1655-1667, nearest.rs:456-482): d_out[i] = d_src[d_idx[i]], zero where the
 *      validity bit of i is clear (d_valid_bits may be NULL = all valid).  elem_bytes in {1, 2, 4, 8, 16}. */
BRG_API int brg_take_dev(brg_ctx* ctx, const void* d_src, uint32_t elem_bytes, const uint32_t* d_idx,
                 const uint32_t* d_valid_bits, uint64_t n, void* d_out, void* stream);

/* ---- variable-length results ------------------------------------------------------------------------ */
BRG_API uint64_t brg_rows_len(const brg_rows* r);
BRG_API const uint32_t* brg_rows_build_idx(const brg_rows* r); /* NULL for SEMI/ANTI                   */
BRG_API const uint32_t* brg_rows_probe_idx(const brg_rows* r);
BRG_API const uint8_t* brg_rows_valid_bits(const brg_rows* r); /* NULL when every row is valid (pairs) */
BRG_API const int64_t* brg_rows_distance(const brg_rows* r);   /* NULL unless nearest with distance    */
BRG_API void brg_rows_release(brg_rows* r);

/* ---- merge (replaces the sweep of MergeStream + the per-contig sort of StreamCollector, merge.rs:260-358,
 *      grouped_stream.rs:48-98; SURVEY 8f rank 3: same sort + scan kernels as the index build).  Rows of one key are
 *      sorted by (start, end); a row joins the current run when start <= run_end + min_dist (strict: <), the boundary a
 *      saturating i64 add.  Result rows (key, run start, run end, n_intervals) come out ordered by key id, then start.
 *      Coordinates are i32 here (the reference sweeps i64): the caller narrows and refuses what does not fit.  Rows with
 *      start > end make the sweep inherently sequential: BRG_ERR_UNSUPPORTED. */
typedef struct brg_merged brg_merged;
BRG_API int brg_merge(brg_ctx* ctx, const uint32_t* key_id, const int32_t* start, const int32_t* end, uint64_t n,
              uint32_t n_keys, int64_t min_dist, int strict, brg_merged** out);
/* cluster (replaces the sweep of ClusterStream, cluster.rs:560-720, and the id offsets of ClusterIdCoordinator::build_offsets,
 * cluster.rs:383-414): same runs as brg_merge, but one result row per INPUT row, ordered by (key, start, end, row):
 * brg_merged_key = original row id, brg_merged_start / _end = extent of the row's run, brg_merged_count = run id,
 * numbered globally in key order (ids of key k start after all runs of keys < k). */
BRG_API int brg_cluster(brg_ctx* ctx, const uint32_t* key_id, const int32_t* start, const int32_t* end, uint64_t n,
                uint32_t n_keys, int64_t min_dist, int strict, brg_merged** out);
BRG_API uint64_t brg_merged_len(const brg_merged* m);
BRG_API const uint32_t* brg_merged_key(const brg_merged* m);
BRG_API const int32_t* brg_merged_start(const brg_merged* m);
BRG_API const int32_t* brg_merged_end(const brg_merged* m);
BRG_API const uint32_t* brg_merged_count(const brg_merged* m);
BRG_API void brg_merged_release(brg_merged* m);

/* ---- subtract (replaces the sweep of SubtractStream, subtract.rs:395-440): the pieces of every left interval [start, end]
 *      that no right interval covers.  `idx` is the index of the RIGHT table built with BRG_WANT_COVERAGE (its merged
 *      intervals are what the sweep walks).  Result rows, left-row-major, pieces of one row in coordinate order:
 *      brg_rows_probe_idx = left row, brg_rows_piece_start / _end = the piece.  strict: a right interval ending AT the left
 *      start / starting AT the left end does not touch it (subtract.rs:399-404, 414-417). */
BRG_API int brg_subtract(brg_ctx* ctx, const brg_index* idx, const uint32_t* key_id, const int32_t* start, const int32_t* end,
                 uint64_t n, int strict, brg_rows** out);
BRG_API const int32_t* brg_rows_piece_start(const brg_rows* r);
BRG_API const int32_t* brg_rows_piece_end(const brg_rows* r);
/* row order of a table sorted by (key, start, end, row): the order in which the sort + sweep table functions walk and emit
 * their input (grouped_stream.rs:92-94, 228-237).  out_row_order: host buffer of n u32. */
BRG_API int brg_sort_rows(brg_ctx* ctx, const uint32_t* key_id, const int32_t* start, const int32_t* end, uint64_t n,
                  uint32_t n_keys, uint32_t* out_row_order);

/* ---- complement (replaces ComplementStream::merge_intervals + emit_contig_complement and the view-only pass,
 *      complement.rs:297-358, 437-457): the gaps that the merged input intervals (run rule of brg_merge with min_dist 0:
 *      start <= run_end joins, strict: <) leave inside each view interval [view_start, view_end).  Views are i64 (the implicit
 *      view of a contig without one is (0, i64::MAX), complement.rs:403-406) and are answered in the order given; a view whose
 *      key is BRG_KEY_ABSENT (or >= n_keys: a contig with no input rows) comes out whole.  Result rows, view-major, gaps
 *      of one view in coordinate order: brg_gaps_view = view row, brg_gaps_start / _end = the gap (half-open as in the
 *      reference: the end of one merged interval to the start of the next). */
typedef struct brg_gaps brg_gaps;
BRG_API int brg_complement(brg_ctx* ctx, const uint32_t* key_id, const int32_t* start, const int32_t* end, uint64_t n,
                   uint32_t n_keys, const uint32_t* view_key, const int64_t* view_start, const int64_t* view_end,
                   uint64_t n_views, int strict, brg_gaps** out);
BRG_API uint64_t brg_gaps_len(const brg_gaps* g);
BRG_API const uint32_t* brg_gaps_view(const brg_gaps* g);
BRG_API const int64_t* brg_gaps_start(const brg_gaps* g);
BRG_API const int64_t* brg_gaps_end(const brg_gaps* g);
BRG_API void brg_gaps_release(brg_gaps* g);

/* ---- instrumentation: kernels launched by this library on the calling process (bench `gpu_launches`) -- */
BRG_API uint64_t brg_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* BIORANGES_H */
