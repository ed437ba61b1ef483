// hark_internal.cuh — shared host/device internals of libhark.so (sm_100a only).
//
// Data model in HBM: a table is n rows x m columns stored as m separate, 256-byte aligned,
// device-resident column arrays (SoA), one dtype per column.  The reference keeps ONE row-major
// 2-D host array and re-copies it across the FFI on every query (table.py:28,
// FutharkContext.py:65,70); here the transpose to SoA happens once at upload.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string.h>

#include <map>
#include <unordered_map>
#include <string>
#include <vector>

#include "../../include/hark.h"

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct hark_col {
    void *ptr = nullptr;
    int32_t dtype = HARK_I32;
    bool owned = true;
    // column statistics (zone map): min / max order key, computed by the first operator that needs them and kept
    // with the column (columns are immutable once built)
    mutable bool mm_valid = false;
    mutable uint64_t mm_lo = 0, mm_hi = 0;
    // f32 columns: exponent of the highest / lowest set bit over all non-zero values, and whether every value is finite
    // (dense_agg.cu: proves when a SUM can be accumulated EXACTLY in fixed point)
    mutable bool fx_valid = false;
    mutable int fx_hi = 0, fx_lo = 0;
    mutable bool fx_finite = false, fx_any = false;
};

struct hark_table {
    int64_t n = 0;   // rows
    int64_t cap = 0; // rows each owned column was allocated for (>= n)
    std::vector<hark_col> cols;
};

struct hark_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;
    cudaMemPool_t pool = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    uint64_t *h_scalars = nullptr; // pinned, 256 x u64: device->host result scalars
    std::string err;
    bool has_err = false;
    hark_stats last = {};
    bool stats_pending = false; // events recorded, elapsed not yet read
    int64_t total_launches = 0;
    int64_t entry_launches = 0;
    int entry_depth = 0; // operators call each other (HAVING -> K1, multi-key GROUP BY -> GROUP BY): only the outermost begin/end count
    std::map<std::string, int64_t> opts;     // tuning knobs (hark_context_set_option)
    std::map<std::string, int64_t> counters; // read-only facts about the last operation ("sort.last_passes", ...)
    struct hk_peer_state *peer = nullptr; // K8c: receive arena + the other ranks' arenas (repartition.cu)

    int fail(int code, const std::string &msg) {
        err = msg;
        has_err = true;
        return code;
    }
    int64_t opt(const char *key, int64_t dflt) const {
        auto it = opts.find(key);
        return it == opts.end() ? dflt : it->second;
    }
    // stream-ordered allocation from the context's pool (never returns memory to the OS until
    // the context dies: repeated queries do not pay cudaMalloc/cudaFree).  On top of the driver's pool sits a cache of
    // freed blocks by size (r2): a block of a size seen before is handed out again without a driver call.  The
    // driver's own reuse re-maps physical pages whenever a request fits no free range exactly, and that showed up as
    // 13 ms for a 2 GB block and 27 ms for an 8 GB one, every query, in the NCCL exchange path
    // (profiles/r02_ai_alloc_trace.txt).  Safe because every kernel that touches pool memory runs on `stream`, in host
    // program order (the upload / download staging on copy_stream is synchronised before its buffers are freed).
    int dalloc(void **p, size_t bytes);
    void dfree(void *p);
    void release_cached_blocks();          // hand every cached block back to the driver's pool
    std::unordered_map<void *, size_t> live_blocks;   // size of every block handed out
    std::multimap<size_t, void *> free_blocks;         // cached blocks by size
    size_t cached_bytes = 0;
    void count_launch(int n = 1) {
        total_launches += n;
        entry_launches += n;
    }
    void entry_begin();                    // reset per-entry stats, record ev_t0
    void entry_end(int64_t alg_bytes, int64_t rows_in, int64_t rows_out);
    // kernel_ms spans the entry's first marked kernel to its last (nested operators extend the span, never restart it)
    bool kernel_marked = false;
    void kernel_begin() {
        if (!kernel_marked) cudaEventRecord(ev_k0, stream);
        kernel_marked = true;
    }
    void kernel_end() { cudaEventRecord(ev_k1, stream); }
};

#define HK_STR2(x) #x
#define HK_STR(x) HK_STR2(x)
#define HK_CUDA(ctx, call)                                                                                 \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return (ctx)->fail(e__ == cudaErrorMemoryAllocation ? HARK_ERR_OOM : HARK_ERR_CUDA,            \
                               std::string(__FILE__ ":" HK_STR(__LINE__) ": " #call ": ") +                \
                                   cudaGetErrorString(e__));                                               \
    } while (0)
#define HK_CHECK_LAUNCH(ctx) HK_CUDA(ctx, cudaGetLastError())
#define HK_TRY(expr)                \
    do {                            \
        int rc__ = (expr);          \
        if (rc__ != HARK_OK) return rc__; \
    } while (0)
#define HK_ARG(ctx, cond, msg)                                 \
    do {                                                       \
        if (!(cond)) return (ctx)->fail(HARK_ERR_ARG, (msg));  \
    } while (0)

// Wraps an ABI entry so that no C++ exception crosses the C boundary.
#define HK_ABI_BEGIN try {
#define HK_ABI_END(ctx)                                                         \
    }                                                                           \
    catch (const std::bad_alloc &) {                                            \
        return (ctx) ? (ctx)->fail(HARK_ERR_OOM, "host allocation failed") : HARK_ERR_OOM; \
    }                                                                           \
    catch (const std::exception &e) {                                           \
        return (ctx) ? (ctx)->fail(HARK_ERR_ARG, e.what()) : HARK_ERR_ARG;      \
    }                                                                           \
    catch (...) {                                                               \
        return (ctx) ? (ctx)->fail(HARK_ERR_ARG, "unknown C++ exception") : HARK_ERR_ARG; \
    }

static inline int hk_dtype_size(int dt) { return (dt == HARK_I64 || dt == HARK_F64) ? 8 : 4; }
static inline bool hk_dtype_ok(int dt) { return dt >= HARK_I32 && dt <= HARK_F64; }
static inline bool hk_dtype_int(int dt) { return dt == HARK_I32 || dt == HARK_U32 || dt == HARK_I64; }

void hk_peer_destroy(hark_ctx *ctx); // repartition.cu

// frees a table from INSIDE an operator: unlike the ABI's hark_table_free it does not pass through HK_ENTER, which
// would reset entry_depth (and with it the outer entry's statistics) in the middle of a nested operator
static inline void hk_table_free(hark_ctx *ctx, hark_table *t) {
    if (!t) return;
    for (auto &c : t->cols)
        if (c.owned) ctx->dfree(c.ptr);
    delete t;
}

// new table with m owned columns of capacity cap rows (uninitialised)
int hk_table_alloc(hark_ctx *ctx, hark_table **out, int64_t n, int64_t cap, const int32_t *dtypes, int64_t m);

// ---- operators implemented in the other translation units ----
int hk_filter(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *cols, int64_t k,
              const hark_pred *preds, int64_t np);
int hk_copy_columns(hark_ctx *ctx, hark_table *dst, const hark_table *src, const int32_t *cols, int64_t k,
                    int64_t src_row0, int64_t nrows, int64_t dst_row0);

int hk_groupby(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t g_col, const int32_t *s_cols,
               const int32_t *ops, int64_t c, const hark_pred *having, int64_t nh, bool pinned_u32);
int hk_groupby_multi(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *g_cols, int64_t ng,
                     const int32_t *s_cols, const int32_t *ops, int64_t c, const hark_pred *having, int64_t nh);
int hk_orderby(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *cols, int64_t k,
               const int32_t *key_cols, const int32_t *desc, int64_t nk);
int hk_join(hark_ctx *ctx, hark_table **out, const hark_table *db1, const hark_table *db2, int32_t col1, int32_t col2,
            const int32_t *cols1, int64_t l, const int32_t *cols2, int64_t k);
int hk_join_ex(hark_ctx *ctx, hark_table **out, const hark_table *db1, const hark_table *db2, int32_t col1, int32_t col2,
               const int32_t *cols1, int64_t l, const int32_t *cols2, int64_t k, int32_t order, bool pinned_u32);
int hk_join_groupby(hark_ctx *ctx, hark_table **out, const hark_table *fact, const hark_table *dim, int32_t fk_col,
                    int32_t pk_col, int32_t g_col, const int32_t *s_cols, const int32_t *ops, int64_t c);
int hk_partition_by_hash(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t key_col, int32_t nparts,
                         int64_t *counts_out);

struct hk_sort_key {  // one radix-sort key column
    const void *ptr;  // column data (not moved unless it is also in the carried set)
    int32_t dtype;
    int32_t desc;
};

// ------------------------------------------------------------------------------------------
// device side helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

#define HK_FULL_MASK 0xffffffffu

__device__ __forceinline__ uint64_t hk_ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void hk_st_relaxed_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t hk_ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void hk_st_relaxed_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint64_t hk_warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(HK_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ uint32_t hk_warp_sum_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(HK_FULL_MASK, v, o);
    return v;
}
// inclusive warp scan
__device__ __forceinline__ uint32_t hk_warp_incl_scan_u32(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(HK_FULL_MASK, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ---- single-pass chained scan ("decoupled look-back") over tiles ----
// state word = 2 flag bits (0 not ready, 1 tile aggregate, 2 inclusive prefix) | 62-bit value.
// Flag and value share one 8-byte word, so a relaxed load observes a consistent pair.
constexpr uint64_t HK_LB_AGG = 1ull << 62;
constexpr uint64_t HK_LB_INC = 2ull << 62;
constexpr uint64_t HK_LB_VAL = (1ull << 62) - 1;

// Called by ALL 32 lanes of one warp.  Publishes `aggregate` for `tile`, walks predecessors
// 32 at a time until an inclusive prefix is found, publishes this tile's inclusive prefix and
// returns the exclusive prefix (same value in every lane).  Tiles must be claimed in increasing
// order by resident CTAs (ticket counter), which makes the wait deadlock-free.
__device__ __forceinline__ uint64_t hk_lookback_u64(uint64_t *state, int64_t tile, uint64_t aggregate) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) hk_st_relaxed_u64(&state[0], HK_LB_INC | aggregate);
        return 0;
    }
    if (lane == 0) hk_st_relaxed_u64(&state[tile], HK_LB_AGG | aggregate);
    uint64_t excl = 0;
    int64_t idx = tile - 1 - lane;
    while (true) {
        uint64_t v = (idx >= 0) ? hk_ld_relaxed_u64(&state[idx]) : HK_LB_INC;
        while (__any_sync(HK_FULL_MASK, (v >> 62) == 0)) {
            if ((v >> 62) == 0) v = hk_ld_relaxed_u64(&state[idx]);
        }
        const unsigned inc_mask = __ballot_sync(HK_FULL_MASK, (v >> 62) == 2);
        const uint64_t val = v & HK_LB_VAL;
        if (inc_mask) {
            const int first = __ffs(inc_mask) - 1; // nearest predecessor holding an inclusive prefix
            excl += hk_warp_sum_u64(lane <= first ? val : 0ull);
            break;
        }
        excl += hk_warp_sum_u64(val);
        idx -= 32;
    }
    if (lane == 0) hk_st_relaxed_u64(&state[tile], HK_LB_INC | (excl + aggregate));
    return excl;
}

// ---- generator (the test oracle restates it bit for bit) ----
__host__ __device__ __forceinline__ uint64_t hk_mix64(uint64_t seed, uint64_t col, uint64_t row) {
    uint64_t z = (seed ^ ((col + 1) * 0xD6E8FEB86659FD93ULL)) + (row + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// ---- order-preserving key transforms (DESIGN.md §sort keys) ----
// signed ints: flip the sign bit; floats: IEEE flip, every NaN maps to the maximum key (NaN last).
__device__ __forceinline__ uint32_t hk_ordkey32(uint32_t bits, int dtype) {
    if (dtype == HARK_U32) return bits;
    if (dtype == HARK_I32) return bits ^ 0x80000000u;
    if ((bits & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu; // NaN
    return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
}
__device__ __forceinline__ uint64_t hk_ordkey64(uint64_t bits, int dtype) {
    if (dtype == HARK_I64) return bits ^ 0x8000000000000000ull;
    if ((bits & 0x7fffffffffffffffull) > 0x7ff0000000000000ull) return ~0ull; // NaN
    return (bits & 0x8000000000000000ull) ? ~bits : (bits | 0x8000000000000000ull);
}

#endif // __CUDACC__
