// dense_agg.cuh — internal interface of K2 (partitioned shared-memory aggregation) and of the fast
// single-pass range partition it is built on (dense_agg.cu).
#pragma once
#include <type_traits>

#include "hark_internal.cuh"

#define HK_DENSE_MAX_VALS 4
#define HK_DENSE_MAX_AGGS 16

// digit of a row = (ordkey(raw) - base) < span ? ((ordkey(raw) - base) >> shift) : 0
struct hk_part_spec {
    int32_t dtype;  // hark_dtype of the key column (defines ordkey)
    uint64_t base;  // smallest order key
    uint64_t span;  // number of distinct order keys covered (max - base + 1); 0 = 2^64
    int shift;
    int nbins;      // <= 256
};

// One partition pass: rows are moved so that bin b occupies [offsets[b], offsets[b+1]); order inside a bin is
// unspecified (the callers aggregate, which is order-independent).  Key 4 or 8 bytes, nv <= 3 carried arrays
// of 4 bytes each.  key_out / vals_out / d_offsets are fresh pool buffers owned by the caller.
int hk_partition_pass(hark_ctx *ctx, int64_t n, const void *key, int kw, const hk_part_spec &spec, int nv,
                      const void *const *vals, void **key_out, void **vals_out, unsigned long long **d_offsets);

// K8t: tile-local partition (partition.cu).  rows: [num_tiles * HK_TPART_TILE][rw] u32 words, row = [key words | value
// words]; dir[t * nbins + b] = start | end << 16 of bin b's run inside tile t.  hash_mask != 0: the bin is a slice of
// a hash table, (hk_hash_key(raw) & hash_mask) >> spec.shift, instead of a key range.
#define HK_TPART_TILE 4096
struct hk_tpart {
    uint32_t *rows = nullptr;
    uint32_t *dir = nullptr;
    int64_t num_tiles = 0;
    int nbins = 0;
    int rw = 0;
};
int hk_tile_partition(hark_ctx *ctx, int64_t n, const void *key, int kw, const hk_part_spec &spec, uint64_t hash_mask, int nv,
                      const void *const *vals, hk_tpart *out);

#ifdef __CUDACC__
// hash of a join key: the slot of the open-addressing tables in join.cu and the slice K8t partitions by
template <int KW>
__device__ __forceinline__ uint64_t hk_hash_key(typename std::conditional<KW == 4, uint32_t, uint64_t>::type raw) {
    uint64_t z = (uint64_t)raw * 0x9E3779B97F4A7C15ULL;
    z ^= z >> 32;
    z *= 0xD6E8FEB86659FD93ULL;
    return z ^ (z >> 29);
}
#endif

struct hk_dense_req {
    int64_t n = 0;
    const void *key = nullptr;  // group key column, or the fact foreign-key column when lut != nullptr
    int32_t key_dtype = HARK_I32;
    int32_t out_key_dtype = HARK_I32;  // dtype of the output key column (group key dtype; U32 for the pinned entry)
    uint64_t g_lo = 0, g_hi = 0;       // order-key range of the GROUP key (R = g_hi - g_lo + 1 dense slots)
    // join mode: group slot + 1 = lut[fk - pk_min] (0 = no match); rows with fk outside [pk_min, pk_min+pk_span) miss
    const uint32_t *lut = nullptr;
    long long pk_min = 0, pk_span = 0;
    // hash join mode: group slot + 1 from an open-addressing table over hk_hash_key(fk) & hmask (join.cu builds it):
    // 8-byte entries {key, slot + 1} for 4-byte keys, 16-byte entries {key, slot + 1 | dim row << 32} for 8-byte keys
    const void *htab = nullptr;
    uint64_t hmask = 0;
    int nvals = 0;
    const void *vals[HK_DENSE_MAX_VALS] = {nullptr, nullptr, nullptr, nullptr};
    int32_t val_dtypes[HK_DENSE_MAX_VALS] = {0, 0, 0, 0};
    const hark_col *val_cols[HK_DENSE_MAX_VALS] = {nullptr, nullptr, nullptr, nullptr}; // the table columns (statistics cache), or null
    int c = 0;                         // output aggregates
    int agg_val[HK_DENSE_MAX_AGGS];    // index into vals, -1 for COUNT
    int agg_code[HK_DENSE_MAX_AGGS];   // hark_agg (already normalised: unknown -> MIN for the pinned entry)
    bool pinned_u32 = false;
};

// Tries the dense path.  *handled = false (and HARK_OK) when the request is not eligible (key range too wide,
// 8-byte value columns, ...): the caller then takes the sort path.
int hk_dense_groupby(hark_ctx *ctx, hark_table **out, const hk_dense_req &rq, bool *handled);

// min / max order key of a raw column (one streaming read)
int hk_col_minmax(hark_ctx *ctx, const void *col, int32_t dtype, int64_t n, uint64_t *lo, uint64_t *hi);
// same for a table column, through the column's cached statistics when `dtype` is the column's own dtype
int hk_column_minmax(hark_ctx *ctx, const hark_col &col, int64_t n, int32_t dtype, uint64_t *lo, uint64_t *hi);
