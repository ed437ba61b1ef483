/*
 * hark_futhark_compat.h — link-compatible aliases for the C API that `futhark c --library
 * futhark/main.fut -o main` (setup.sh:12) generates, so that a program (or the CFFI module
 * `build_futhark_ffi main` produces, setup.sh:13) written against the generated main.h can link
 * libhark.so instead.  The generated header is not in the reference tree (it is a build product);
 * the names and signatures below follow Futhark's documented C API for the entry points of
 * main.fut:7,9 (`query_sel`, `query_groupby`) and for join.fut:52 (`join`, an entry the
 * reference never compiles) with the array types those entries use (i32 1-d, i32 2-d, u32 2-d).
 *
 * Semantics kept: int results, 0 = success; futhark_new_* copies host data in and returns an
 * owned opaque handle (NULL on failure); entries borrow inputs and return fresh outputs;
 * futhark_context_get_error returns a malloc'd string the caller frees (NULL if none).
 */
#ifndef HARK_FUTHARK_COMPAT_H
#define HARK_FUTHARK_COMPAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct futhark_context_config;
struct futhark_context;
struct futhark_i32_1d;
struct futhark_i32_2d;
struct futhark_u32_2d;

struct futhark_context_config *futhark_context_config_new(void);
void futhark_context_config_free(struct futhark_context_config *cfg);
void futhark_context_config_set_debugging(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_profiling(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_logging(struct futhark_context_config *cfg, int flag);
/* extension: which CUDA device the context drives (default: the current device) */
void futhark_context_config_set_device(struct futhark_context_config *cfg, int device);

struct futhark_context *futhark_context_new(struct futhark_context_config *cfg);
void futhark_context_free(struct futhark_context *ctx);
int futhark_context_sync(struct futhark_context *ctx);
char *futhark_context_get_error(struct futhark_context *ctx);
int futhark_context_clear_caches(struct futhark_context *ctx);

struct futhark_i32_1d *futhark_new_i32_1d(struct futhark_context *ctx, const int32_t *data, int64_t dim0);
int futhark_free_i32_1d(struct futhark_context *ctx, struct futhark_i32_1d *arr);
int futhark_values_i32_1d(struct futhark_context *ctx, struct futhark_i32_1d *arr, int32_t *data);
const int64_t *futhark_shape_i32_1d(struct futhark_context *ctx, struct futhark_i32_1d *arr);

struct futhark_i32_2d *futhark_new_i32_2d(struct futhark_context *ctx, const int32_t *data, int64_t dim0, int64_t dim1);
int futhark_free_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr);
int futhark_values_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr, int32_t *data);
const int64_t *futhark_shape_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr);

struct futhark_u32_2d *futhark_new_u32_2d(struct futhark_context *ctx, const uint32_t *data, int64_t dim0, int64_t dim1);
int futhark_free_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr);
int futhark_values_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr, uint32_t *data);
const int64_t *futhark_shape_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr);

/* main.fut:7 */
int futhark_entry_query_sel(struct futhark_context *ctx, struct futhark_i32_2d **out0, const struct futhark_i32_2d *in0,
                            const struct futhark_i32_1d *in1);
/* main.fut:9 */
int futhark_entry_query_groupby(struct futhark_context *ctx, struct futhark_u32_2d **out0,
                                const struct futhark_u32_2d *in0, const int32_t in1, const struct futhark_i32_1d *in2,
                                const struct futhark_i32_1d *in3);
/* join.fut:52 */
int futhark_entry_join(struct futhark_context *ctx, struct futhark_u32_2d **out0, const struct futhark_u32_2d *in0,
                       const struct futhark_u32_2d *in1, const int32_t in2, const int32_t in3,
                       const struct futhark_i32_1d *in4, const struct futhark_i32_1d *in5);

#ifdef __cplusplus
}
#endif
#endif /* HARK_FUTHARK_COMPAT_H */
