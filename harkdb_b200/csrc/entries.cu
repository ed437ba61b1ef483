// entries.cu — the C-ABI query entries (argument checking + dispatch to the operator kernels).
//
// hark_entry_query_sel      <- main.fut:7  query_sel      (select.fut:9-23)
// hark_entry_query_groupby  <- main.fut:9  query_groupby  (groupby.fut:51-62)
// hark_entry_join           <- join.fut:52 join           (orphan entry in the reference)
// The *_ex / filter / orderby / join_groupby entries are extensions.
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "hark_internal.cuh"

#define HK_ENTER(ctx)                \
    if (!(ctx)) return HARK_ERR_ARG; \
    (ctx)->entry_depth = 0;          \
    HK_CUDA(ctx, cudaSetDevice((ctx)->device))

extern "C" int hark_entry_query_sel(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *cols,
                                    int64_t k) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db && k >= 0 && (k == 0 || cols), "query_sel: bad argument");
    return hk_filter(ctx, out, db, cols, k, nullptr, 0);
    HK_ABI_END(ctx)
}

extern "C" int hark_entry_query_filter(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *cols,
                                       int64_t k, const hark_pred *preds, int64_t np) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db && k >= 0 && (k == 0 || cols) && np >= 0 && (np == 0 || preds), "query_filter: bad argument");
    return hk_filter(ctx, out, db, cols, k, preds, np);
    HK_ABI_END(ctx)
}

extern "C" int hark_entry_query_groupby(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t g_col,
                                        const int32_t *s_cols, const int32_t *t_cols, int64_t c) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db && c >= 0 && (c == 0 || (s_cols && t_cols)), "query_groupby: bad argument");
    return hk_groupby(ctx, out, db, g_col, s_cols, t_cols, c, nullptr, 0, /*pinned_u32=*/true);
    HK_ABI_END(ctx)
}

extern "C" int hark_entry_query_groupby_ex(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t g_col,
                                           const int32_t *s_cols, const int32_t *ops, int64_t c,
                                           const hark_pred *having, int64_t nh) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db && c >= 0 && (c == 0 || (s_cols && ops)) && nh >= 0 && (nh == 0 || having),
           "query_groupby_ex: bad argument");
    return hk_groupby(ctx, out, db, g_col, s_cols, ops, c, having, nh, /*pinned_u32=*/false);
    HK_ABI_END(ctx)
}

extern "C" int hark_entry_query_groupby_multi(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *g_cols,
                                              int64_t ng, const int32_t *s_cols, const int32_t *ops, int64_t c,
                                              const hark_pred *having, int64_t nh) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db && g_cols && ng >= 1 && c >= 0 && (c == 0 || (s_cols && ops)) && nh >= 0 && (nh == 0 || having),
           "query_groupby_multi: bad argument");
    if (ng == 1) { // one key: the single-key operator, with the key column first as everywhere else
        return hk_groupby(ctx, out, db, g_cols[0], s_cols, ops, c, having, nh, /*pinned_u32=*/false);
    }
    return hk_groupby_multi(ctx, out, db, g_cols, ng, s_cols, ops, c, having, nh);
    HK_ABI_END(ctx)
}

extern "C" int hark_entry_query_orderby(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *cols,
                                        int64_t k, const int32_t *key_cols, const int32_t *desc, int64_t nk) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db && k >= 0 && (k == 0 || cols) && nk >= 0 && (nk == 0 || key_cols),
           "query_orderby: bad argument");
    return hk_orderby(ctx, out, db, cols, k, key_cols, desc, nk);
    HK_ABI_END(ctx)
}

extern "C" int hark_entry_join(hark_ctx *ctx, hark_table **out, const hark_table *db1, const hark_table *db2,
                               int32_t col1, int32_t col2, const int32_t *cols1, int64_t l, const int32_t *cols2,
                               int64_t k) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db1 && db2 && l >= 0 && k >= 0 && (l == 0 || cols1) && (k == 0 || cols2), "join: bad argument");
    return hk_join(ctx, out, db1, db2, col1, col2, cols1, l, cols2, k);
    HK_ABI_END(ctx)
}

extern "C" int hark_entry_join_ex(hark_ctx *ctx, hark_table **out, const hark_table *db1, const hark_table *db2,
                                  int32_t col1, int32_t col2, const int32_t *cols1, int64_t l, const int32_t *cols2,
                                  int64_t k, int32_t order) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db1 && db2 && l >= 0 && k >= 0 && (l == 0 || cols1) && (k == 0 || cols2), "join_ex: bad argument");
    return hk_join_ex(ctx, out, db1, db2, col1, col2, cols1, l, cols2, k, order, false);
    HK_ABI_END(ctx)
}

extern "C" int hark_entry_join_groupby(hark_ctx *ctx, hark_table **out, const hark_table *fact, const hark_table *dim,
                                       int32_t fk_col, int32_t pk_col, int32_t g_col, const int32_t *s_cols,
                                       const int32_t *ops, int64_t c) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && fact && dim && c >= 0 && (c == 0 || (s_cols && ops)), "join_groupby: bad argument");
    return hk_join_groupby(ctx, out, fact, dim, fk_col, pk_col, g_col, s_cols, ops, c);
    HK_ABI_END(ctx)
}

extern "C" int hark_table_sort_by(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t key_col) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db, "table_sort_by: bad argument");
    const int64_t m = (int64_t)db->cols.size();
    std::vector<int32_t> all;
    for (int64_t c = 0; c < m; c++) all.push_back((int32_t)c);
    int32_t zero = 0;
    return hk_orderby(ctx, out, db, all.data(), m, &key_col, &zero, 1);
    HK_ABI_END(ctx)
}

extern "C" int hark_table_partition_by_hash(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t key_col,
                                            int32_t nparts, int64_t *counts_out) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db && counts_out && nparts >= 1 && nparts <= 256, "partition_by_hash: bad argument");
    return hk_partition_by_hash(ctx, out, db, key_col, nparts, counts_out);
    HK_ABI_END(ctx)
}
