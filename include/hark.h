/*
 * hark.h — C ABI of libhark.so, the B200-native replacement for the Futhark-generated
 * library behind HarkDB's relational operator path.
 *
 * What it replaces (reference file:line, all under /root/reference):
 *   - the CFFI module `_main` built by setup.sh:12-13 (`futhark c --library futhark/main.fut`
 *     + `build_futhark_ffi main`) and driven from FutharkContext.py:41,65-66,70-71;
 *   - entry points main.fut:7 (`query_sel`) and main.fut:9 (`query_groupby`), plus the
 *     orphan entry join.fut:52 (`join`), with the argument order of those entries kept;
 *   - the array marshalling futhark_ffi does around them (futhark_new_* / futhark_values_* /
 *     futhark_shape_* / futhark_free_*): here hark_table_from_host / hark_table_to_host /
 *     hark_table_shape / hark_table_free.  A literally link-compatible `futhark_*` alias
 *     layer is declared in include/hark_futhark_compat.h.
 *
 * Conventions (same as the Futhark C API): every function returns int, 0 = success; inputs
 * are borrowed and never consumed; outputs are fresh handles owned by the caller; no C++
 * exception or abort() crosses this boundary for a user-level fault; a context serialises
 * its calls (one call at a time per context); the last error text is retrievable with
 * hark_context_get_error().  One context drives ONE GPU (one process per GPU; the
 * multi-GPU layer lives above this ABI).
 *
 * There is no CPU fallback: every entry runs hand-written sm_100a kernels and fails with
 * HARK_ERR_CUDA when no device is usable.
 */
#ifndef HARK_H
#define HARK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HARK_ABI_VERSION 1

typedef struct hark_ctx hark_ctx;     /* opaque: device, stream, memory pool, scratch, stats */
typedef struct hark_table hark_table; /* opaque: n rows x m device-resident columns (SoA)    */

typedef enum { HARK_I32 = 0, HARK_U32 = 1, HARK_I64 = 2, HARK_F32 = 3, HARK_F64 = 4 } hark_dtype;

typedef enum {
    HARK_OK = 0,
    HARK_ERR_ARG = 1,         /* bad argument, incl. a column index out of bounds (the Futhark
                                 entry would return non-zero with "Index [i] out of bounds") */
    HARK_ERR_CUDA = 2,        /* CUDA runtime / launch failure, or no usable device            */
    HARK_ERR_OOM = 3,         /* device or host allocation failed                              */
    HARK_ERR_UNSUPPORTED = 4  /* well-formed request this build does not implement             */
} hark_status;

/* Comparison operators of a WHERE / HAVING conjunct. */
typedef enum { HARK_GT = 0, HARK_GE = 1, HARK_LT = 2, HARK_LE = 3, HARK_EQ = 4, HARK_NE = 5 } hark_cmp;

/* Flags OR-ed into hark_pred.op.  A predicate list is in conjunctive normal form: HARK_PRED_OR chains a predicate
 * with the NEXT one into an OR-clause (a maximal chain; the last predicate of the list must not carry it), the
 * result is the AND of the clauses; HARK_PRED_NOT negates the comparison's boolean result (so NOT (x > c) is true
 * for a NaN x, unlike x <= c).  A list without flags is the plain conjunction.                                    */
#define HARK_PRED_OP_MASK 0xff
#define HARK_PRED_OR 0x100
#define HARK_PRED_NOT 0x200

/* One conjunct `column <op> constant`.  Integer columns (i32/u32/i64) are widened to int64
 * and compared with `ival`; f32 columns compare in f32 against (float)fval, f64 columns in
 * f64 against fval; NaN compares false except under HARK_NE (IEEE).                        */
typedef struct {
    int32_t col;   /* column index into the table the predicate is evaluated on */
    int32_t op;    /* hark_cmp */
    int64_t ival;
    double fval;
} hark_pred;

/* Aggregate codes.  0-4 are exactly parse.py:81 / groupby.fut:35-41 (0 and any unknown code
 * fall through to `min`, groupby.fut:41); 5, 6 and 7 are extensions.                         */
typedef enum {
    HARK_AGG_KEY = 0, HARK_AGG_PROD = 1, HARK_AGG_SUM = 2, HARK_AGG_MAX = 3, HARK_AGG_MIN = 4,
    HARK_AGG_COUNT = 5, HARK_AGG_AVG = 6,
    HARK_AGG_SUMF64 = 7, /* f64 sum as its own f64 column: the partial aggregate behind a distributed AVG */
    HARK_AGG_SUM64 = 8   /* exact SUM: integer columns accumulate in 64 bits -> i64 column (SQL's SUM; code 2 keeps the
                            reference's wrap-around in the column width, groupby.fut:37); float columns: as code 7 */
} hark_agg;

/* Synthetic column generator (hark_table_synth).  Value of (column c, global row r) is a pure
 * function of (seed, c, r): h = hark_mix64(seed, c, r) as defined in DESIGN.md §generator, so
 * the host oracle can regenerate any row range.                                              */
typedef enum {
    HARK_GEN_UNIFORM = 0, /* ints: lo + mulhi64(h, range)  (range 0: lo + h, wrapping);
                             f32: fmaf(u24, fhi-flo, flo), u24 = (h>>40)*2^-24;
                             f64: fma (u53, fhi-flo, flo), u53 = (h>>11)*2^-53               */
    HARK_GEN_AFFINE = 1,  /* ((a*r + b) mod 2^64) mod range (range 0: no mod) — unique keys when
                             gcd(a, range) = 1 and a*r+b does not wrap                       */
    HARK_GEN_CONST = 2,   /* lo (ints) / flo (floats)                                         */
    HARK_GEN_AFFINE_UNIFORM = 4, /* a * mulhi64(h, range) + b (wrapping, truncated to the dtype): a uniform draw from the key
                             set {a*j + b : 0 <= j < range} — the foreign keys of a dimension whose primary key is
                             HARK_GEN_AFFINE with the same a, b and range 0 (sparse keys when a is large and odd)      */
    HARK_GEN_LOGUNIFORM = 3 /* skewed keys, integer arithmetic only: octave e = mulhi64(h, floor(log2 range))
                             uniform, then uniform inside the octave: k = 2^e + (h2 & (2^e - 1)) - 1, value =
                             lo + k mod range, h2 = hark_mix64(seed ^ 0x5851F42D4C957F2D, c, r).  P(k) ~ 1/(k+1):
                             a piecewise-constant Zipf(1.0) — key 0 holds 1/floor(log2 range) of all rows       */
} hark_gen_kind;

typedef struct {
    int32_t kind; /* hark_gen_kind */
    int32_t reserved;
    int64_t lo;
    uint64_t range;
    double flo, fhi;
    uint64_t a, b;
} hark_colspec;

/* Per-entry measurements taken with CUDA events on the context's stream. */
typedef struct {
    double kernel_ms;   /* device time of the entry's dominant kernel(s)                     */
    double total_ms;    /* device time of the whole entry (first launch to last)             */
    int64_t alg_bytes;  /* algorithmic bytes of the entry              */
    int64_t rows_in;
    int64_t rows_out;
    int64_t launches;   /* kernels this entry launched                                       */
} hark_stats;

/* ---- context (replaces futhark_context_config_new / futhark_context_new, FutharkContext.py:41) ---- */
int hark_abi_version(void);
/* device: CUDA ordinal, -1 = the calling thread's current device.
 * stream: a cudaStream_t to launch on (borrowed; cudaStreamLegacy / cudaStreamPerThread are accepted), or NULL
 *         for a non-blocking stream the context owns.                                                            */
hark_ctx *hark_context_new(int device, void *stream);
void hark_context_free(hark_ctx *ctx);
int hark_context_sync(hark_ctx *ctx);
/* The context's stream-ordered memory pool keeps freed blocks for the next query (no cudaMalloc in steady state), and a
 * cache of freed blocks by size sits on top of it (blocks <= "pool.cache_block_gb" = 12, at most "pool.cache_gb" = 48 in
 * total; "pool.cache" = 0 switches it off): memory in that cache is not visible to other allocators in the process.  This
 * call empties the cache and hands every free block back to the driver (after a synchronize) — for callers that switch
 * between workloads of very different sizes on a nearly full device, or that need the memory for their own tensors.      */
int hark_context_trim(hark_ctx *ctx);
char *hark_context_get_error(hark_ctx *ctx); /* malloc'd, caller frees; NULL if no error      */
int hark_context_device(hark_ctx *ctx);
/* Text of the last failure of hark_context_new (static storage), for when it returned NULL.  */
const char *hark_last_init_error(void);

/* ---- tables (replace futhark_new_{i32,u32}_2d / futhark_values_* / futhark_shape_* / futhark_free_*) ---- */
/* Row-major homogeneous host array [n][m] -> device SoA (table.py:52-74 data model).          */
int hark_table_from_host(hark_ctx *ctx, hark_table **out, const void *rowmajor, int64_t n, int64_t m,
                         int32_t dtype);
/* m host column arrays, one dtype each. */
int hark_table_from_columns(hark_ctx *ctx, hark_table **out, const void *const *host_cols,
                            const int32_t *dtypes, int64_t n, int64_t m);
/* m device column arrays, borrowed: the table never frees them.  Contract for borrowed columns: (1) 16-byte aligned
 * and readable up to the next multiple of 256 bytes past the last row (the kernels read ragged ends with 128-bit
 * loads; cudaMalloc / torch allocations satisfy this); (2) the caller may change the buffers BETWEEN entries — the
 * library keeps no statistics (min / max, zone maps) of borrowed columns, only of columns it owns — but not while an
 * entry is running on the context's stream.                                                                       */
int hark_table_from_device(hark_ctx *ctx, hark_table **out, void *const *dev_cols, const int32_t *dtypes,
                           int64_t n, int64_t m);
/* Generated on the device; rows are global rows row0 .. row0+n-1 of the synthetic relation.    */
int hark_table_synth(hark_ctx *ctx, hark_table **out, int64_t n, int64_t m, const int32_t *dtypes,
                     uint64_t seed, const hark_colspec *specs, int64_t row0);
int hark_table_shape(hark_ctx *ctx, const hark_table *t, int64_t shape[2]);
int hark_table_dtypes(hark_ctx *ctx, const hark_table *t, int32_t *dtypes_out /* [m] */);
/* Device SoA -> row-major host [n][m]; all columns must share one dtype.                      */
int hark_table_to_host(hark_ctx *ctx, const hark_table *t, void *rowmajor_out);
int hark_table_column_to_host(hark_ctx *ctx, const hark_table *t, int32_t col, int64_t row0, int64_t nrows,
                              void *out);
void *hark_table_column_ptr(hark_ctx *ctx, const hark_table *t, int32_t col); /* device pointer, borrowed */
int hark_table_free(hark_ctx *ctx, hark_table *t);

/* ---- reference-pinned entries: argument order of main.fut:7,9 and join.fut:52-54 ---- */
/* out[r][j] = db[r][cols[j]]  (select.fut:9-23).  Any dtype; row order kept.                   */
int hark_entry_query_sel(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *cols, int64_t k);
/* groupby.fut:51-62: all columns i32/u32, compared and combined as u32; one row per distinct
 * key ascending unsigned; row = [key, agg_1..agg_c], agg_i = fold(op t_cols[i-1]) over
 * db[rows of group][s_cols[i-1]], codes 0-4 as hark_agg (0/unknown -> min).  Output u32.       */
int hark_entry_query_groupby(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t g_col,
                             const int32_t *s_cols, const int32_t *t_cols, int64_t c);
/* join.fut:52-75: inner equi-join on db1[:,col1] == db2[:,col2] (u32); row = db1[r1][cols1] ++
 * db2[r2][cols2]; ordered by key ascending unsigned, then r1, then r2.                         */
int hark_entry_join(hark_ctx *ctx, hark_table **out, const hark_table *db1, const hark_table *db2, int32_t col1,
                    int32_t col2, const int32_t *cols1, int64_t l, const int32_t *cols2, int64_t k);

/* ---- extensions (no reference implementation; semantics in DESIGN.md §extensions) ---- */
/* SELECT cols WHERE p_1 AND ... AND p_np; input row order kept (what `filter` at select.fut:18
 * would do).                                                                                   */
int hark_entry_query_filter(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *cols, int64_t k,
                            const hark_pred *preds, int64_t np);
/* Typed GROUP BY: key column any integer dtype, ordered by its own signedness; value columns any
 * dtype; codes 0-8; SUM/PROD wrap in the column's width (SUM64 does not), COUNT -> i64 column, AVG -> f64 column;
 * HAVING = conjunction over OUTPUT column indices (0 = key).                                   */
int hark_entry_query_groupby_ex(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t g_col,
                                const int32_t *s_cols, const int32_t *ops, int64_t c, const hark_pred *having,
                                int64_t nh);
/* GROUP BY several integer key columns — what parse.py:64 ("TODO: allow to be several columns") asks for.  Output
 * columns: the ng keys in g_cols order, then the c aggregates (codes and types as hark_entry_query_groupby_ex);
 * rows ascending lexicographically by the keys, each in its own signedness.  HAVING indexes the output columns.
 * The keys' combined value ranges (max - min per key) must fit 63 bits, else HARK_ERR_UNSUPPORTED.              */
int hark_entry_query_groupby_multi(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *g_cols,
                                   int64_t ng, const int32_t *s_cols, const int32_t *ops, int64_t c,
                                   const hark_pred *having, int64_t nh);
/* SELECT cols ORDER BY key_cols[0] [DESC], key_cols[1] ... ; stable w.r.t. input row order;
 * signed order for i32/i64, IEEE total order with NaN last for f32/f64.                       */
int hark_entry_query_orderby(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *cols, int64_t k,
                             const int32_t *key_cols, const int32_t *desc, int64_t nk);
/* SELECT d.g_col, agg(f.s_cols) FROM fact f JOIN dim d ON f.fk_col = d.pk_col GROUP BY d.g_col
 * (dim.pk_col unique).  Output as hark_entry_query_groupby_ex.                                */
int hark_entry_join_groupby(hark_ctx *ctx, hark_table **out, const hark_table *fact, const hark_table *dim,
                            int32_t fk_col, int32_t pk_col, int32_t g_col, const int32_t *s_cols, const int32_t *ops,
                            int64_t c);

/* Typed inner equi-join: the key columns are any integer dtype (the same on both sides, compared in that dtype's own
 * order), the projected columns keep their dtypes.
 *   order = 1: rows ordered as hark_entry_join does (join.fut:55-75: key ascending, then db1 row, then db2 row) — both
 *              sides sorted (K3) and merged; db1 and db2 below 2^32-1 rows;
 *   order = 0: hash build on db2 + probe with db1: rows grouped by db1 row, ascending, the matches of one db1 row in
 *              unspecified order (a multiset result); db1 of any size, db2 below 2^32-1 rows.  When no key occurs
 *              twice in db2 the join is one pass and its result columns are ALLOCATED for db1's row count (every probe
 *              row can match once); the table reports the true row count.  "join.hash_one_pass" = 0 takes the
 *              count-then-expand plan, which allocates exactly.                                                      */
int hark_entry_join_ex(hark_ctx *ctx, hark_table **out, const hark_table *db1, const hark_table *db2, int32_t col1,
                       int32_t col2, const int32_t *cols1, int64_t l, const int32_t *cols2, int64_t k, int32_t order);

/* ---- building blocks exported for the multi-GPU layer and for tests ---- */
/* Stable LSD radix sort of whole rows by one column (ascending; signedness of the dtype).      */
int hark_table_sort_by(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t key_col);
/* Stable partition of rows into nparts (<= 256) buckets by key RANGE: bucket of a row = number of splitters
 * that are <= its key tuple (lexicographic over key_cols, each compared through the ORDER BY order key:
 * signed ints by sign-bit flip, floats IEEE with NaN last, DESC columns complemented).  splitters is a host
 * array [nparts-1][nk] of such order keys (uint64), ascending.  Rows of bucket p are contiguous, in input order.
 * This is what ORDER BY / GROUP BY / JOIN repartition with across GPUs (sampled splitters).                 */
int hark_table_partition_by_splitters(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *key_cols,
                                      const int32_t *desc, int64_t nk, const uint64_t *splitters, int32_t nparts,
                                      int64_t *counts_out);
/* Order keys (uint64, the same mapping as above) of the given rows: out[i][j] = ordkey(db[rows[i]][key_cols[j]]).
 * Used to sample splitter candidates.                                                                       */
int hark_table_sample_order_keys(hark_ctx *ctx, const hark_table *db, const int32_t *key_cols, const int32_t *desc,
                                 int64_t nk, const int64_t *rows, int64_t nrows, uint64_t *out);
/* Last step of a distributed GROUP BY: `merged` holds [key, partial columns...] where every aggregate of `ops`
 * contributed its partial columns in order (AVG: an f64 sum and an i64 count; everything else one column).
 * Output [key, agg_1..agg_c] with AVG = sum / count.                                                       */
int hark_entry_groupby_finalize(hark_ctx *ctx, hark_table **out, const hark_table *merged, const int32_t *ops,
                                int64_t c);
/* ---- K8c: partition fused with the exchange over NVLink peer memory (one process per GPU, CUDA IPC) ----
 * Every rank creates a receive arena and passes its 64-byte IPC handle around (any transport; the Python layer
 * uses torch.distributed); after hark_peer_arena_open the scatter below stores rows directly into the destination
 * GPUs' arenas.  Protocol per exchange: count -> (ranks all-gather counts) -> run -> (one stream-ordered
 * collective) -> result.  The result table is a view of this rank's arena, valid until the next exchange.       */
int hark_peer_arena_create(hark_ctx *ctx, int64_t bytes, void *ipc_handle_out /* 64 bytes */);
int hark_peer_arena_open(hark_ctx *ctx, const void *handles /* world x 64 bytes, rank order */, int32_t world,
                         int32_t my_rank);
int hark_peer_arena_close(hark_ctx *ctx);
/* arena bytes a rank needs to receive `rows` rows of db's schema */
int64_t hark_peer_arena_bytes_needed(hark_ctx *ctx, const hark_table *db, int64_t rows);
/* destination of a row = number of splitters <= its key tuple (as hark_table_partition_by_splitters, nparts = world) */
int hark_peer_scatter_count(hark_ctx *ctx, const hark_table *db, const int32_t *key_cols, const int32_t *desc, int64_t nk,
                            const uint64_t *splitters, int64_t *counts_out /* [world] */);
int hark_peer_scatter_run(hark_ctx *ctx, const hark_table *db, const int64_t *counts_matrix /* [src][dst], world x world */);
int hark_peer_scatter_result(hark_ctx *ctx, hark_table **out);

/* Stable partition of rows into nparts buckets by mix(key) % nparts (integer key column);
 * counts_out[nparts] (host) receives the bucket sizes; rows of bucket p are contiguous.        */
int hark_table_partition_by_hash(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t key_col,
                                 int32_t nparts, int64_t *counts_out);
/* Rows [row0, row0+nrows) of t as a new table (device copy). */
int hark_table_slice(hark_ctx *ctx, hark_table **out, const hark_table *t, int64_t row0, int64_t nrows);
/* Concatenate two tables with equal schemas. */
int hark_table_concat(hark_ctx *ctx, hark_table **out, const hark_table *a, const hark_table *b);

/* ---- measurement ---- */
int hark_stats_last(hark_ctx *ctx, hark_stats *out);
int64_t hark_stats_total_launches(hark_ctx *ctx);
/* Tuning knobs for experiments ("filter.impl", "filter.ctas_per_sm", "dense.dynamic", "sort.straddle", "join.build",
 * "pool.cache", ...: the list is `known[]` in csrc/context.cu, every one is described where DESIGN.md discusses its
 * kernel); returns HARK_ERR_ARG for an unknown key.                                             */
int hark_context_set_option(hark_ctx *ctx, const char *key, int64_t value);
/* Reads back an option, or one of the read-only counters the last sort left behind: "sort.last_passes",
 * "sort.last_truncated" (1: only the top digits were sorted and ties repaired), "sort.last_fix_runs",
 * "sort.last_fallback" (1: the repair met a long run and the sort was redone with every pass).
 * HARK_ERR_ARG if the key was never set.                                                        */
int hark_context_get_option(hark_ctx *ctx, const char *key, int64_t *value);

/* The sort's pass-truncation decision as a pure host function, for tests without a device: bits[k] =
 * significant bits of key k (most significant first), sample = S normalised key tuples (row-major).  Returns 1 and
 * fills the outputs when the sort would run only the top digits (keys 0..kstar, key kstar by its top q digits, its
 * bits below `shift` left to the tie repair), 0 when it keeps every pass.                                         */
int hark_debug_plan_truncation(int64_t n, int32_t nk, const int32_t *bits, const uint64_t *sample, int32_t S,
                               int32_t slack, int32_t *kstar, int32_t *q, int32_t *shift, int32_t *passes);

/* ---- pinned host memory for callers that want DMA-speed uploads ---- */
void *hark_host_alloc(int64_t bytes);
void hark_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* HARK_H */
