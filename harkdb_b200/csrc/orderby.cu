// orderby.cu — ORDER BY (multi-key, ASC/DESC, stable) and hash partitioning on top of the radix sort.
//
// The reference has no ORDER BY operator (README.md:15 lists "Sort By"; nothing implements it); its only
// sort is the one inside groupby.fut / join.fut.  Semantics here are oracle-defined:
// lexicographic over the key columns, signed order for i32/i64, IEEE order with NaN last for floats, DESC =
// exact reverse key order, ties keep input row order.
//
// Data movement: only the key columns (and, when other columns are selected, one row-id column) ride through
// the radix passes; the remaining selected columns are gathered once at the end (K7).
#include <algorithm>
#include <new>
#include <stdexcept>
#include <vector>

#include "hark_internal.cuh"
#include "dense_agg.cuh"
#include "sort.cuh"

namespace {

struct OutBuilder { // assembles an owned result table from raw device buffers
    hark_ctx *ctx;
    hark_table *t;
    explicit OutBuilder(hark_ctx *c, int64_t n) : ctx(c), t(new hark_table()) {
        t->n = n;
        t->cap = n;
    }
    void push_owned(void *ptr, int32_t dtype) {
        hark_col c;
        c.ptr = ptr;
        c.dtype = dtype;
        c.owned = true;
        t->cols.push_back(c);
    }
    void abandon() {
        if (!t) return;
        for (auto &c : t->cols) ctx->dfree(c.ptr);
        delete t;
        t = nullptr;
    }
};

} // namespace

int hk_orderby(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *cols, int64_t k,
               const int32_t *key_cols, const int32_t *desc, int64_t nk) {
    const int64_t n = db->n, m = (int64_t)db->cols.size();
    for (int64_t j = 0; j < k; j++) HK_ARG(ctx, cols[j] >= 0 && cols[j] < m, "query_orderby: selected column index out of bounds");
    for (int64_t j = 0; j < nk; j++) HK_ARG(ctx, key_cols[j] >= 0 && key_cols[j] < m, "query_orderby: key column index out of bounds");
    if (nk == 0) return hk_filter(ctx, out, db, cols, k, nullptr, 0);

    ctx->entry_begin();
    // distinct key columns, in priority order (a repeated key column can never break a tie)
    std::vector<hk_sort_keyspec> keys;
    std::vector<hk_sort_array> arrays;
    std::vector<int> key_array_of_col((size_t)m, -1);
    for (int64_t j = 0; j < nk; j++) {
        const int c = key_cols[j];
        if (key_array_of_col[c] >= 0) continue;
        hk_sort_array a;
        a.in = db->cols[c].ptr;
        a.width = hk_dtype_size(db->cols[c].dtype);
        key_array_of_col[c] = (int)arrays.size();
        arrays.push_back(a);
        hk_sort_keyspec ks;
        ks.array = key_array_of_col[c];
        ks.dtype = db->cols[c].dtype;
        ks.desc = desc ? (desc[j] != 0) : 0;
        if (n > 0) { // the column's cached min / max order key saves the sort a pass over the column
            HK_TRY(hk_column_minmax(ctx, db->cols[c], n, ks.dtype, &ks.lo, &ks.hi));
            ks.have_range = true;
        }
        keys.push_back(ks);
    }
    bool need_rowid = false;
    for (int64_t j = 0; j < k; j++)
        if (key_array_of_col[cols[j]] < 0) need_rowid = true;
    HK_ARG(ctx, (int)arrays.size() + (need_rowid ? 1 : 0) <= HK_SORT_MAX_ARRAYS, "query_orderby: too many key columns");

    void *rowid_in = nullptr;
    int rowid_idx = -1;
    const int rw = n > 0xffffffffll ? 8 : 4;
    if (need_rowid) {
        HK_TRY(ctx->dalloc(&rowid_in, (size_t)std::max<int64_t>(n, 1) * rw));
        int rc = hk_iota(ctx, rowid_in, n, rw);
        if (rc != HARK_OK) {
            ctx->dfree(rowid_in);
            return rc;
        }
        hk_sort_array a;
        a.in = rowid_in;
        a.width = rw;
        rowid_idx = (int)arrays.size();
        arrays.push_back(a);
    }
    hk_sort_info info;
    int rc = hk_radix_sort(ctx, n, keys, arrays, 0, nullptr, &info);
    ctx->dfree(rowid_in);
    if (rc != HARK_OK) return rc;

    OutBuilder ob(ctx, n);
    std::vector<bool> taken(arrays.size(), false);
    for (int64_t j = 0; j < k && rc == HARK_OK; j++) {
        const int c = cols[j];
        const int32_t dt = db->cols[c].dtype;
        const int w = hk_dtype_size(dt);
        const int ai = key_array_of_col[c];
        if (ai >= 0 && !taken[ai]) { // the sorted key array becomes the output column
            ob.push_owned(arrays[ai].result, dt);
            taken[ai] = true;
            continue;
        }
        void *p = nullptr;
        rc = ctx->dalloc(&p, (size_t)std::max<int64_t>(n, 1) * w);
        if (rc != HARK_OK) break;
        ob.push_owned(p, dt);
        if (ai >= 0) rc = hk_copy_bytes(ctx, p, arrays[ai].result, n * w);
        else rc = hk_gather(ctx, p, db->cols[c].ptr, w, arrays[rowid_idx].result, rw, n);
    }
    for (size_t a = 0; a < arrays.size(); a++)
        if (!taken[a]) ctx->dfree(arrays[a].result);
    if (rc != HARK_OK) {
        ob.abandon();
        return rc;
    }
    int64_t alg = 0;
    std::vector<int> seen((size_t)m, 0);
    for (int64_t j = 0; j < nk; j++) seen[key_cols[j]] = 1;
    for (int64_t j = 0; j < k; j++) seen[cols[j]] = 1;
    for (int64_t c = 0; c < m; c++)
        if (seen[c]) alg += n * hk_dtype_size(db->cols[c].dtype);
    for (int64_t j = 0; j < k; j++) alg += n * hk_dtype_size(db->cols[cols[j]].dtype);
    ctx->entry_end(alg, n, n);
    ctx->last.launches = ctx->entry_launches;
    *out = ob.t;
    return HARK_OK;
}

int hk_partition_by_hash(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t key_col, int32_t nparts,
                         int64_t *counts_out) {
    const int64_t n = db->n, m = (int64_t)db->cols.size();
    HK_ARG(ctx, key_col >= 0 && key_col < m, "partition_by_hash: key column index out of bounds");
    HK_ARG(ctx, hk_dtype_int(db->cols[key_col].dtype), "partition_by_hash: key column must be an integer column");
    ctx->entry_begin();
    // carried: the key column + a row id when the table is wide, else every column directly
    const bool direct = m <= HK_SORT_MAX_ARRAYS;
    std::vector<hk_sort_array> arrays;
    std::vector<hk_sort_keyspec> keys;
    void *rowid_in = nullptr;
    const int rw = n > 0xffffffffll ? 8 : 4;
    if (direct) {
        for (int64_t c = 0; c < m; c++) {
            hk_sort_array a;
            a.in = db->cols[c].ptr;
            a.width = hk_dtype_size(db->cols[c].dtype);
            arrays.push_back(a);
        }
        keys.push_back(hk_sort_keyspec{key_col, db->cols[key_col].dtype, 0});
    } else {
        hk_sort_array a;
        a.in = db->cols[key_col].ptr;
        a.width = hk_dtype_size(db->cols[key_col].dtype);
        arrays.push_back(a);
        HK_TRY(ctx->dalloc(&rowid_in, (size_t)std::max<int64_t>(n, 1) * rw));
        int rc = hk_iota(ctx, rowid_in, n, rw);
        if (rc != HARK_OK) {
            ctx->dfree(rowid_in);
            return rc;
        }
        hk_sort_array r;
        r.in = rowid_in;
        r.width = rw;
        arrays.push_back(r);
        keys.push_back(hk_sort_keyspec{0, db->cols[key_col].dtype, 0});
    }
    int rc = hk_radix_sort(ctx, n, keys, arrays, nparts, counts_out, nullptr);
    ctx->dfree(rowid_in);
    if (rc != HARK_OK) return rc;
    OutBuilder ob(ctx, n);
    if (direct) {
        for (int64_t c = 0; c < m; c++) ob.push_owned(arrays[c].result, db->cols[c].dtype);
    } else {
        for (int64_t c = 0; c < m && rc == HARK_OK; c++) {
            const int w = hk_dtype_size(db->cols[c].dtype);
            void *p = nullptr;
            rc = ctx->dalloc(&p, (size_t)std::max<int64_t>(n, 1) * w);
            if (rc != HARK_OK) break;
            ob.push_owned(p, db->cols[c].dtype);
            rc = hk_gather(ctx, p, db->cols[c].ptr, w, arrays[1].result, rw, n);
        }
        ctx->dfree(arrays[0].result);
        ctx->dfree(arrays[1].result);
        if (rc != HARK_OK) {
            ob.abandon();
            return rc;
        }
    }
    int64_t alg = 0;
    for (int64_t c = 0; c < m; c++) alg += 2 * n * hk_dtype_size(db->cols[c].dtype);
    ctx->entry_end(alg, n, n);
    *out = ob.t;
    return HARK_OK;
}
