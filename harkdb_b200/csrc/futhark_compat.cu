// futhark_compat.cu — the `futhark_*` alias layer declared in include/hark_futhark_compat.h:
// thin wrappers that give libhark.so the symbol names and calling conventions of the C library
// `futhark c --library futhark/main.fut` generates (setup.sh:12), over the hark_* ABI.
#include <stdlib.h>

#include <new>
#include <vector>

#include "../../include/hark_futhark_compat.h"
#include "hark_internal.cuh"

struct futhark_context_config {
    int device = -1;
    int debugging = 0, profiling = 0, logging = 0;
};
struct futhark_context {
    hark_ctx *h = nullptr;
};
struct futhark_i32_1d {
    std::vector<int32_t> v;
    int64_t shape[1];
};
struct fut_2d {
    hark_table *t = nullptr;
    int64_t shape[2];
};
struct futhark_i32_2d : fut_2d {};
struct futhark_u32_2d : fut_2d {};

extern "C" {

struct futhark_context_config *futhark_context_config_new(void) { return new (std::nothrow) futhark_context_config(); }
void futhark_context_config_free(struct futhark_context_config *cfg) { delete cfg; }
void futhark_context_config_set_debugging(struct futhark_context_config *cfg, int flag) { if (cfg) cfg->debugging = flag; }
void futhark_context_config_set_profiling(struct futhark_context_config *cfg, int flag) { if (cfg) cfg->profiling = flag; }
void futhark_context_config_set_logging(struct futhark_context_config *cfg, int flag) { if (cfg) cfg->logging = flag; }
void futhark_context_config_set_device(struct futhark_context_config *cfg, int device) { if (cfg) cfg->device = device; }

struct futhark_context *futhark_context_new(struct futhark_context_config *cfg) {
    hark_ctx *h = hark_context_new(cfg ? cfg->device : -1, nullptr);
    if (!h) return nullptr;
    futhark_context *c = new (std::nothrow) futhark_context();
    if (!c) {
        hark_context_free(h);
        return nullptr;
    }
    c->h = h;
    return c;
}
void futhark_context_free(struct futhark_context *ctx) {
    if (!ctx) return;
    hark_context_free(ctx->h);
    delete ctx;
}
int futhark_context_sync(struct futhark_context *ctx) { return ctx ? hark_context_sync(ctx->h) : 1; }
char *futhark_context_get_error(struct futhark_context *ctx) { return ctx ? hark_context_get_error(ctx->h) : nullptr; }
int futhark_context_clear_caches(struct futhark_context *ctx) { return ctx ? 0 : 1; }

struct futhark_i32_1d *futhark_new_i32_1d(struct futhark_context *ctx, const int32_t *data, int64_t dim0) {
    if (!ctx || dim0 < 0 || (dim0 > 0 && !data)) return nullptr;
    futhark_i32_1d *a = new (std::nothrow) futhark_i32_1d();
    if (!a) return nullptr;
    try {
        a->v.assign(data, data + dim0);
    } catch (...) {
        delete a;
        return nullptr;
    }
    a->shape[0] = dim0;
    return a;
}
int futhark_free_i32_1d(struct futhark_context *, struct futhark_i32_1d *arr) {
    delete arr;
    return 0;
}
int futhark_values_i32_1d(struct futhark_context *, struct futhark_i32_1d *arr, int32_t *data) {
    if (!arr || (!data && !arr->v.empty())) return 1;
    for (size_t i = 0; i < arr->v.size(); i++) data[i] = arr->v[i];
    return 0;
}
const int64_t *futhark_shape_i32_1d(struct futhark_context *, struct futhark_i32_1d *arr) { return arr ? arr->shape : nullptr; }

} // extern "C"

template <typename A>
static A *new_2d(struct futhark_context *ctx, const void *data, int64_t d0, int64_t d1, int dtype) {
    if (!ctx) return nullptr;
    A *a = new (std::nothrow) A();
    if (!a) return nullptr;
    if (hark_table_from_host(ctx->h, &a->t, data, d0, d1, dtype) != HARK_OK) {
        delete a;
        return nullptr;
    }
    a->shape[0] = d0;
    a->shape[1] = d1;
    return a;
}
static int free_2d(struct futhark_context *ctx, fut_2d *a) {
    if (!a) return 0;
    int rc = ctx ? hark_table_free(ctx->h, a->t) : 1;
    delete a;
    return rc;
}
template <typename A>
static int wrap_out(struct futhark_context *ctx, A **out, hark_table *t) {
    A *a = new (std::nothrow) A();
    if (!a) {
        hk_table_free(ctx->h, t);
        return HARK_ERR_OOM;
    }
    a->t = t;
    hark_table_shape(ctx->h, t, a->shape);
    *out = a;
    return 0;
}

extern "C" {

struct futhark_i32_2d *futhark_new_i32_2d(struct futhark_context *ctx, const int32_t *data, int64_t dim0, int64_t dim1) {
    return new_2d<futhark_i32_2d>(ctx, data, dim0, dim1, HARK_I32);
}
int futhark_free_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr) { return free_2d(ctx, arr); }
int futhark_values_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr, int32_t *data) {
    return (ctx && arr) ? hark_table_to_host(ctx->h, arr->t, data) : 1;
}
const int64_t *futhark_shape_i32_2d(struct futhark_context *, struct futhark_i32_2d *arr) { return arr ? arr->shape : nullptr; }

struct futhark_u32_2d *futhark_new_u32_2d(struct futhark_context *ctx, const uint32_t *data, int64_t dim0, int64_t dim1) {
    return new_2d<futhark_u32_2d>(ctx, data, dim0, dim1, HARK_U32);
}
int futhark_free_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr) { return free_2d(ctx, arr); }
int futhark_values_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr, uint32_t *data) {
    return (ctx && arr) ? hark_table_to_host(ctx->h, arr->t, data) : 1;
}
const int64_t *futhark_shape_u32_2d(struct futhark_context *, struct futhark_u32_2d *arr) { return arr ? arr->shape : nullptr; }

int futhark_entry_query_sel(struct futhark_context *ctx, struct futhark_i32_2d **out0, const struct futhark_i32_2d *in0,
                            const struct futhark_i32_1d *in1) {
    if (!ctx || !out0 || !in0 || !in1) return 1;
    hark_table *t = nullptr;
    int rc = hark_entry_query_sel(ctx->h, &t, in0->t, in1->v.data(), (int64_t)in1->v.size());
    return rc ? rc : wrap_out(ctx, out0, t);
}

int futhark_entry_query_groupby(struct futhark_context *ctx, struct futhark_u32_2d **out0,
                                const struct futhark_u32_2d *in0, const int32_t in1, const struct futhark_i32_1d *in2,
                                const struct futhark_i32_1d *in3) {
    if (!ctx || !out0 || !in0 || !in2 || !in3) return 1;
    if (in3->v.size() < in2->v.size()) return 1; // groupby.fut:47 would index t_cols out of bounds
    hark_table *t = nullptr;
    int rc = hark_entry_query_groupby(ctx->h, &t, in0->t, in1, in2->v.data(), in3->v.data(), (int64_t)in2->v.size());
    return rc ? rc : wrap_out(ctx, out0, t);
}

int futhark_entry_join(struct futhark_context *ctx, struct futhark_u32_2d **out0, const struct futhark_u32_2d *in0,
                       const struct futhark_u32_2d *in1, const int32_t in2, const int32_t in3,
                       const struct futhark_i32_1d *in4, const struct futhark_i32_1d *in5) {
    if (!ctx || !out0 || !in0 || !in1 || !in4 || !in5) return 1;
    hark_table *t = nullptr;
    int rc = hark_entry_join(ctx->h, &t, in0->t, in1->t, in2, in3, in4->v.data(), (int64_t)in4->v.size(), in5->v.data(),
                             (int64_t)in5->v.size());
    return rc ? rc : wrap_out(ctx, out0, t);
}

} // extern "C"
