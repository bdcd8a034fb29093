/*
 * bioranges_host.h -- host-side mirror of datafusion-bio-function-ranges' operator surface, above the C ABI of
 * bioranges.h.  The reference is a Rust DataFusion extension; this image has no Rust toolchain, so the operators that
 * call the interval-overlap hot path are restated in C++ (csrc/host/brg_host.cpp) with the same names, argument
 * meaning, output schemas and error behaviour, over the Arrow C Data Interface (what arrow-rs' FFI_ArrowArray /
 * FFI_ArrowSchema and pyarrow's _export_to_c / _import_from_c speak).  Paths below are relative to
 * /root/reference/datafusion/bio-function-ranges/src.
 *
 *   brgh_session                 create_bio_session() + BioConfig / `SET bio.*`      session_context.rs:16-148
 *   brgh_register_batch          ctx.register_batch / MemTable (the tests' fixtures)  tests/integration_test.rs:38-60
 *   brgh_count_overlaps/coverage CountOverlapsProvider + get_count_stream/get_stream  count_overlaps.rs:46-169,
 *                                                                                     interval_tree.rs:155-280
 *   brgh_nearest                 NearestProvider + get_nearest_stream                 nearest.rs:44-547
 *   brgh_overlap                 OverlapProvider (overlap() UDTF)                     overlap.rs:70-330
 *   brgh_interval_join           IntervalJoinExec (range join rewritten by the planner rule)
 *                                                                                     physical_planner/joins/interval_join.rs
 *   column resolution + errors   get_join_col_arrays / PosArray::resolve              array_utils.rs:73-242
 *
 * Every function returns 0 or a status; the message of the last failure on the calling thread is brg_last_error()
 * (bioranges.h).  Statuses >= 10 name the DataFusionError variant the reference raises at the same point.
 * Arrow ownership follows the C Data Interface: arrays/schemas passed IN are moved (their release pointer is cleared),
 * arrays/schemas handed OUT are released by the caller.
 */
#ifndef BIORANGES_HOST_H
#define BIORANGES_HOST_H

#include "bioranges.h"

#ifdef __cplusplus
extern "C" {
#endif

struct ArrowArray;
struct ArrowSchema;

enum {
    BRGH_ERR_PLAN = 10,            /* DataFusionError::Plan            */
    BRGH_ERR_EXECUTION = 11,       /* DataFusionError::Execution       */
    BRGH_ERR_NOT_IMPLEMENTED = 12, /* DataFusionError::NotImplemented  */
    BRGH_ERR_INTERNAL = 13,        /* DataFusionError::Internal        */
    BRGH_ERR_EXTERNAL = 14,        /* DataFusionError::External (ParseAlgorithmError, session_context.rs:74-110) */
    BRGH_ERR_ARROW = 15            /* DataFusionError::ArrowError      */
};

/* join types with dedicated logic in IntervalJoinExec (interval_join.rs:1014-1024); every other type yields Inner rows */
enum { BRGH_JOIN_INNER = 0, BRGH_JOIN_RIGHT_SEMI = 1, BRGH_JOIN_RIGHT_ANTI = 2 };
/* overlap() output modes (table_function.rs:112-160, overlap.rs:24-40) */
enum { BRGH_OVERLAP_JOIN = 0, BRGH_OVERLAP_LEFT = 1, BRGH_OVERLAP_LEFT_DISTINCT = 2 };

typedef struct brgh_session brgh_session;
typedef struct brgh_result brgh_result; /* the record batches a table function / join produced, in stream order */

/* ---- session + configuration -------------------------------------------------------------------------------------- */
/* create_bio_session(): the GPU context is created lazily by the first operator that needs it, so configuration and
 * schema errors surface without a device. */
BRG_API int brgh_create_bio_session(int device_id, brgh_session** out);
BRG_API void brgh_session_destroy(brgh_session* s);
/* `SET bio.<option> = <value>`: prefer_interval_join (bool), interval_join_algorithm (Algorithm, parsed
 * case-insensitively), interval_join_low_memory (bool).  Unknown Algorithm -> BRGH_ERR_EXTERNAL "Can't parse '<v>' as
 * Algorithm". */
BRG_API int brgh_session_set(brgh_session* s, const char* key, const char* value);
/* SHOW bio.<option>: canonical text (Algorithm by its Display name, session_context.rs:92-105) */
BRG_API int brgh_session_get(const brgh_session* s, const char* key, char* buf, uint64_t cap);
/* soft target for coalescing probe batches before a PCIe round trip (DataFusion's 8192-row batches are far too small) */
BRG_API int brgh_session_set_coalesce_rows(brgh_session* s, uint64_t rows);

/* ---- tables --------------------------------------------------------------------------------------------------------- */
/* appends one record batch (a struct array) to the named in-memory table; batch and schema are moved */
BRG_API int brgh_register_batch(brgh_session* s, const char* table, struct ArrowArray* batch, struct ArrowSchema* schema);
BRG_API int brgh_deregister_table(brgh_session* s, const char* table);

/* ---- table functions (register_ranges_functions, table_function.rs:835-895) ------------------------------------------
 * args mirrors the SQL argument list AFTER the two table names: 0, 3 or 6 column names, optionally followed by
 * 'strict' / 'weak' (and for overlap() an output mode) -- parse_col_args_extra / parse_overlap_args,
 * table_function.rs:59-189.  Default columns: contig, pos_start, pos_end.
 * count_overlaps / coverage: LEFT is built, RIGHT is probed; output = right columns + `count` | `coverage` Int64 NOT NULL,
 * one batch per probe batch, rows in probe order (count_overlaps.rs:60-66, interval_tree.rs:269-273). */
BRG_API int brgh_count_overlaps(brgh_session* s, const char* left, const char* right, const char* const* args, int n_args,
                        brgh_result** out);
BRG_API int brgh_coverage(brgh_session* s, const char* left, const char* right, const char* const* args, int n_args,
                  brgh_result** out);
/* nearest(left, right [,k] [,overlap] [,compute_distance] [,cols...] [,'strict']): output = left_* (nullable), right_*,
 * optional `distance` Int64 nullable (nearest.rs:57-78); k >= 1 else BRGH_ERR_PLAN "nearest() k must be >= 1". */
BRG_API int brgh_nearest(brgh_session* s, const char* left, const char* right, int64_t k, int include_overlaps,
                 int compute_distance, const char* const* args, int n_args, brgh_result** out);
/* overlap(left, right, ...): RIGHT is built, LEFT is probed (overlap.rs:173-180) */
BRG_API int brgh_overlap(brgh_session* s, const char* left, const char* right, const char* const* args, int n_args,
                 brgh_result** out);

/* merge(table [, min_dist] [, col1, col2, col3] [, 'strict'|'weak']) (table_function.rs:430-520, merge.rs): output =
 * (<contig col> Utf8, <start col> Int64, <end col> Int64, n_intervals Int64), all NOT NULL, contigs in name order, batches
 * of 8192 rows.  min_dist < 0 -> BRGH_ERR_PLAN.  Coordinates beyond i32 -> BRGH_ERR_NOT_IMPLEMENTED (the reference sweeps
 * i64; the GPU path is 32-bit). */
BRG_API int brgh_merge(brgh_session* s, const char* table, int64_t min_dist, const char* const* args, int n_args,
               brgh_result** out);

/* cluster(table [, min_dist] [, col1, col2, col3] [, 'strict'|'weak']) (table_function.rs, cluster.rs:30-72): every input
 * row plus `cluster`, `cluster_start`, `cluster_end` Int64 NOT NULL; a 3-column input is rebuilt as (Utf8, Int64, Int64), a
 * wider one keeps all its columns; cluster ids are global, numbered in contig-name order (cluster.rs:383-414). */
BRG_API int brgh_cluster(brgh_session* s, const char* table, int64_t min_dist, const char* const* args, int n_args,
                 brgh_result** out);

/* subtract(left, right [, cols...] [, 'strict'|'weak']) (subtract.rs): the pieces of every left interval that no right
 * interval covers, left rows walked in (contig name, start, end, row) order.  A 3-column left table is rebuilt as
 * (Utf8, Int64, Int64); a wider one keeps its columns with start / end replaced by the Int64 pieces (subtract.rs:48-75). */
BRG_API int brgh_subtract(brgh_session* s, const char* left, const char* right, const char* const* args, int n_args,
                  brgh_result** out);

/* complement(table [, view_table] [, col1, col2, col3 [, vcol1, vcol2, vcol3]] [, 'strict'|'weak']) (complement.rs;
 * table_function.rs:602-767): the gaps the merged input intervals leave inside each view interval of their contig (a contig
 * without a view gets (0, i64::MAX)); contigs that only the view table names come out whole, after the others.  `args` are
 * the arguments AFTER the table name, view table included.  Output (contig Utf8, start Int64, end Int64) NOT NULL, named
 * after the input columns. */
BRG_API int brgh_complement(brgh_session* s, const char* table, const char* const* args, int n_args, brgh_result** out);

/* ---- the range join IntervalJoinExec executes ---------------------------------------------------------------------------
 *   SELECT * FROM left l JOIN right r ON l.<on_left[i]> = r.<on_right[i]> ...
 *        AND l.<left_start> <[=] r.<right_end> AND l.<left_end> >[=] r.<right_start>
 * LEFT is the build side (interval_join.rs:69-72).  lt_strict / gt_strict say whether the two comparisons are `<` / `>`:
 * the planner then compares against `end - 1` on that side (intervals.rs:85-115).  Output = left columns ++ right
 * columns (Inner) or the right columns (RightSemi / RightAnti), one batch per probe batch unless
 * bio.interval_join_low_memory caps them (interval_join.rs:1466-1573).  With bio.interval_join_algorithm =
 * CoitreesNearest every probe row yields exactly one row, the left side NULL when the key has no intervals
 * (interval_join.rs:1627-1636). */
BRG_API int brgh_interval_join(brgh_session* s, const char* left, const char* right, const char* const* on_left,
                       const char* const* on_right, int n_on, const char* left_start, const char* left_end,
                       const char* right_start, const char* right_end, int lt_strict, int gt_strict, int join_type,
                       brgh_result** out);
/* The range-join filter parser (physical_planner/intervals.rs:30-218, tests :220-525).  `filter` is the residual of the ON
 * clause as SQL text: the AND of exactly two comparisons (<, <=, >, >=), each between ONE column of the left input and ONE
 * column of the right input, written `alias.column` (optionally quoted, parenthesised or inside CAST(.. AS type)).  The four
 * bounds are assigned by side and direction as parse_condition does; a strict comparison subtracts one from the END it
 * involves (`minus_one`): left_end_minus_one for `l.end > r.start` / `r.start < l.end`, right_end_minus_one for
 * `l.start < r.end` / `r.end > l.start`.  Failures are Internal errors with the reference's texts ("filter expression is not
 * an AND", "Unsupported operator: =", "rs already set", "left end bound not found", "complex sub queries are not supported"). */
typedef struct {
    char left_start[128], left_end[128], right_start[128], right_end[128];
    int left_end_minus_one, right_end_minus_one;
} brgh_col_intervals;
BRG_API int brgh_parse_join_filter(const char* filter, const char* left_alias, const char* right_alias, brgh_col_intervals* out);
/* brgh_interval_join with the range condition given as filter text (what the planner rule extracts from the join,
 * physical_planner/mod.rs + intervals.rs): SELECT * FROM left <left_alias> JOIN right <right_alias> ON <on pairs> AND <filter> */
BRG_API int brgh_interval_join_filter(brgh_session* s, const char* left, const char* right, const char* const* on_left,
                              const char* const* on_right, int n_on, const char* filter, const char* left_alias,
                              const char* right_alias, int join_type, brgh_result** out);
/* text IntervalJoinExec prints in EXPLAIN for this session's algorithm: "IntervalJoinExec: mode=CollectLeft,
 * join_type=Inner, alg=Coitrees" style (interval_join.rs:353-362; the reference's tests assert on `alg=`) */
BRG_API int brgh_explain_interval_join(const brgh_session* s, int join_type, char* buf, uint64_t cap);

/* ---- results ------------------------------------------------------------------------------------------------------------ */
BRG_API uint64_t brgh_result_num_batches(const brgh_result* r);
BRG_API uint64_t brgh_result_num_rows(const brgh_result* r);
/* exports batch i (a struct array) and its schema; both are released by the caller */
BRG_API int brgh_result_export(brgh_result* r, uint64_t i, struct ArrowArray* out, struct ArrowSchema* out_schema);
BRG_API void brgh_result_release(brgh_result* r);

/* ---- pieces usable without a GPU (and unit-tested on the CPU box) ------------------------------------------------------- */
/* PosArray::resolve (array_utils.rs:73-131): resolves column `name` of a record batch to i32; out must hold the
 * batch's row count.  Int32 is copied, Int64/UInt32/UInt64 are narrowed with the reference's overflow check. */
BRG_API int brgh_resolve_positions(struct ArrowArray* batch, struct ArrowSchema* schema, const char* name, const char* label,
                           int32_t* out, uint64_t cap);
/* arrow::compute::take of every column of a record batch by a UInt32 index column with optional validity bitmap
 * (NULL index -> NULL row).  batch/schema are moved; out/out_schema are released by the caller. */
BRG_API int brgh_take(struct ArrowArray* batch, struct ArrowSchema* schema, const uint32_t* idx, const uint8_t* idx_valid_bits,
              uint64_t n, const char* name_prefix, struct ArrowArray* out, struct ArrowSchema* out_schema);

#ifdef __cplusplus
}
#endif
#endif /* BIORANGES_HOST_H */
