// stubs.cu — operators not built yet report HARK_ERR_UNSUPPORTED (never a CPU fallback).
#include "hark_internal.cuh"
int hk_groupby(hark_ctx *ctx, hark_table **, const hark_table *, int32_t, const int32_t *, const int32_t *, int64_t,
               const hark_pred *, int64_t, bool) { return ctx->fail(HARK_ERR_UNSUPPORTED, "groupby: not built yet"); }
int hk_orderby(hark_ctx *ctx, hark_table **, const hark_table *, const int32_t *, int64_t, const int32_t *,
               const int32_t *, int64_t) { return ctx->fail(HARK_ERR_UNSUPPORTED, "orderby: not built yet"); }
int hk_join(hark_ctx *ctx, hark_table **, const hark_table *, const hark_table *, int32_t, int32_t, const int32_t *,
            int64_t, const int32_t *, int64_t) { return ctx->fail(HARK_ERR_UNSUPPORTED, "join: not built yet"); }
int hk_join_groupby(hark_ctx *ctx, hark_table **, const hark_table *, const hark_table *, int32_t, int32_t, int32_t,
                    const int32_t *, const int32_t *, int64_t) { return ctx->fail(HARK_ERR_UNSUPPORTED, "join_groupby: not built yet"); }
int hk_partition_by_hash(hark_ctx *ctx, hark_table **, const hark_table *, int32_t, int32_t, int64_t *) {
    return ctx->fail(HARK_ERR_UNSUPPORTED, "partition_by_hash: not built yet"); }
