// context.cu — context, device-resident tables, host<->device marshalling, synthetic generator.
//
// Replaces what futhark_ffi + the generated C API do around the entries in the reference:
// futhark_context_new (FutharkContext.py:41), futhark_new_*_2d (the per-query copy-in at
// FutharkContext.py:65,70), futhark_values_* / futhark_shape_* (from_futhark, :66,71).
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <stdexcept>

#include "hark_internal.cuh"

static thread_local char g_init_err[512] = "";

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
int hark_ctx::dalloc(void **p, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255; // padded: a 16-byte group load of a ragged column end stays in bounds
    if (bytes == 0) bytes = 256;
    // a cached block of this size, or up to 1/8 larger
    if (opt("pool.cache", 1) != 0) {
        auto it = free_blocks.lower_bound(bytes);
        if (it != free_blocks.end() && it->first <= bytes + bytes / 8) {
            *p = it->second;
            cached_bytes -= it->first;
            live_blocks[*p] = it->first;
            free_blocks.erase(it);
            return HARK_OK;
        }
    }
    static const bool trace_alloc = getenv("HARK_TRACE_ALLOC") != nullptr; // diagnosis: allocations that take > 0.5 ms of host time
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaMallocFromPoolAsync(p, bytes, pool, stream);
    if (trace_alloc) {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 0.5) {
            uint64_t res = 0, used = 0, high = 0;
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &res);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemHigh, &high);
            size_t fr = 0, tot = 0;
            cudaMemGetInfo(&fr, &tot);
            fprintf(stderr, "[hark] dalloc %zu bytes took %.2f ms (device %d): pool reserved %.2f GB (high %.2f), used %.2f GB, device free %.1f GB\n",
                    bytes, ms, device, res / 1073741824.0, high / 1073741824.0, used / 1073741824.0, fr / 1073741824.0);
        }
    }
    if (e == cudaErrorMemoryAllocation) { // out of memory or fragmented: give everything unused back to the driver, retry once
        cudaGetLastError();
        release_cached_blocks();
        cudaStreamSynchronize(stream);
        cudaMemPoolTrimTo(pool, 0);
        e = cudaMallocFromPoolAsync(p, bytes, pool, stream);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        char buf[160];
        snprintf(buf, sizeof buf, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return fail(e == cudaErrorMemoryAllocation ? HARK_ERR_OOM : HARK_ERR_CUDA, buf);
    }
    live_blocks[*p] = bytes;
    return HARK_OK;
}

void hark_ctx::dfree(void *p) {
    if (!p) return;
    auto it = live_blocks.find(p);
    if (it == live_blocks.end()) { // not from dalloc (should not happen): the driver knows what to do
        cudaFreeAsync(p, stream);
        return;
    }
    const size_t bytes = it->second;
    live_blocks.erase(it);
    // Blocks above pool.cache_block_gb (default 12) are not kept: memory in this cache is invisible to the driver, which
    // CAN hand unused pool memory to another allocator in the process (torch's cudaMalloc) when the device runs out —
    // a one-GPU 2e9-row sort frees 32 GB blocks that the caller's own tensors may need next.
    if (opt("pool.cache", 1) == 0 || bytes > ((size_t)opt("pool.cache_block_gb", 12) << 30)) {
        cudaFreeAsync(p, stream);
        return;
    }
    free_blocks.emplace(bytes, p);
    cached_bytes += bytes;
    // bounded: beyond pool.cache_gb (default 48) the largest cached blocks go back to the driver's pool
    const size_t cap = (size_t)opt("pool.cache_gb", 48) << 30;
    while (cached_bytes > cap && !free_blocks.empty()) {
        auto last = std::prev(free_blocks.end());
        cudaFreeAsync(last->second, stream);
        cached_bytes -= last->first;
        free_blocks.erase(last);
    }
}

void hark_ctx::release_cached_blocks() {
    for (auto &kv : free_blocks) cudaFreeAsync(kv.second, stream);
    free_blocks.clear();
    cached_bytes = 0;
}

void hark_ctx::entry_begin() {
    if (entry_depth++ > 0) return; // nested operator: the outer entry's clock keeps running
    entry_launches = 0;
    kernel_marked = false;
    last = hark_stats{};
    cudaEventRecord(ev_t0, stream);
    cudaEventRecord(ev_k0, stream); // entries without a marked kernel report kernel_ms ~ 0
    cudaEventRecord(ev_k1, stream);
}

void hark_ctx::entry_end(int64_t alg_bytes, int64_t rows_in, int64_t rows_out) {
    if (entry_depth > 0 && --entry_depth > 0) return;
    cudaEventRecord(ev_t1, stream);
    last.alg_bytes = alg_bytes;
    last.rows_in = rows_in;
    last.rows_out = rows_out;
    last.launches = entry_launches;
    stats_pending = true;
}

extern "C" int hark_abi_version(void) { return HARK_ABI_VERSION; }

extern "C" const char *hark_last_init_error(void) { return g_init_err; }

extern "C" hark_ctx *hark_context_new(int device, void *stream) {
    g_init_err[0] = 0;
    hark_ctx *ctx = nullptr;
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            snprintf(g_init_err, sizeof g_init_err,
                     "libhark: no usable CUDA device (%s); there is no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
            cudaGetLastError();
            return nullptr;
        }
        if (device < 0) {
            if (cudaGetDevice(&device) != cudaSuccess) device = 0;
        }
        if (device >= ndev) {
            snprintf(g_init_err, sizeof g_init_err, "libhark: device %d out of range (%d devices)", device, ndev);
            return nullptr;
        }
        if ((e = cudaSetDevice(device)) != cudaSuccess) {
            snprintf(g_init_err, sizeof g_init_err, "libhark: cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
            return nullptr;
        }
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, device);
        if (prop.major != 10) {
            snprintf(g_init_err, sizeof g_init_err,
                     "libhark: device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                     prop.minor);
            return nullptr;
        }
        ctx = new hark_ctx();
        ctx->device = device;
        ctx->num_sms = prop.multiProcessorCount;
        if (stream) {
            ctx->stream = (cudaStream_t)stream;
            ctx->own_stream = false;
        } else {
            e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
            ctx->own_stream = true;
        }
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
        cudaMemPoolProps pp;
        memset(&pp, 0, sizeof pp);
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = device;
        if (e == cudaSuccess) e = cudaMemPoolCreate(&ctx->pool, &pp);
        if (e == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            e = cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        cudaEvent_t *evs[] = {&ctx->ev_k0, &ctx->ev_k1, &ctx->ev_t0, &ctx->ev_t1};
        for (auto ev : evs)
            if (e == cudaSuccess) e = cudaEventCreate(ev);
        for (int i = 0; i < 2; i++) {
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&ctx->h_scalars, 256 * sizeof(uint64_t), cudaHostAllocDefault);
        if (e != cudaSuccess) {
            snprintf(g_init_err, sizeof g_init_err, "libhark: context setup failed: %s", cudaGetErrorString(e));
            hark_context_free(ctx);
            return nullptr;
        }
        return ctx;
    } catch (...) {
        snprintf(g_init_err, sizeof g_init_err, "libhark: exception during context creation");
        delete ctx;
        return nullptr;
    }
}

extern "C" void hark_context_free(hark_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->release_cached_blocks();
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    hk_peer_destroy(ctx);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    cudaEvent_t evs[] = {ctx->ev_k0, ctx->ev_k1, ctx->ev_t0, ctx->ev_t1, ctx->ev_copy[0], ctx->ev_copy[1],
                         ctx->ev_done[0], ctx->ev_done[1]};
    for (auto ev : evs)
        if (ev) cudaEventDestroy(ev);
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
}

#define HK_ENTER(ctx)                                       \
    if (!(ctx)) return HARK_ERR_ARG;                        \
    (ctx)->entry_depth = 0;          \
    HK_CUDA(ctx, cudaSetDevice((ctx)->device))

extern "C" int hark_context_sync(hark_ctx *ctx) {
    HK_ENTER(ctx);
    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HARK_OK;
}

extern "C" int hark_context_trim(hark_ctx *ctx) {
    HK_ENTER(ctx);
    ctx->release_cached_blocks();
    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    HK_CUDA(ctx, cudaMemPoolTrimTo(ctx->pool, 0));
    return HARK_OK;
}

extern "C" char *hark_context_get_error(hark_ctx *ctx) {
    if (!ctx || !ctx->has_err) return nullptr;
    char *s = (char *)malloc(ctx->err.size() + 1);
    if (s) memcpy(s, ctx->err.c_str(), ctx->err.size() + 1);
    ctx->has_err = false;
    return s;
}

extern "C" int hark_context_device(hark_ctx *ctx) { return ctx ? ctx->device : -1; }

extern "C" int hark_context_set_option(hark_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) return HARK_ERR_ARG;
    static const char *known[] = {"filter.impl", "filter.ctas_per_sm", "sort.ctas_per_sm", "groupby.impl",
                                  "join.impl",   "upload.chunk_mb",    "dense.log2_slots",  "dense.smem_bytes",
                                  "part.ctas_per_sm", "part.threads", "join.lut_slice_bytes", "sort.impl", "sort.rank", "stats.cache",
                                  "sort.trunc", "sort.trunc_slack", "dense.part_impl", "join.build", "dense.fixed_point",
                                  "sort.digit_bits", "dense.spec", "sort.fuse2", "sort.sweep16", "sort.sweep16_min_rows", "join.build_partition", "join.build_partition_min_rows", "sort.straddle", "sort.sweep16_fast", "sort.fix_fast", "dense.dynamic", "pool.cache", "pool.cache_gb", "pool.cache_block_gb", "join.carry", "join.hash_one_pass", "dense.home_by_rows", "tpart.ctas_per_sm", nullptr};
    for (int i = 0; known[i]; i++)
        if (!strcmp(known[i], key)) {
            ctx->opts[key] = value;
            return HARK_OK;
        }
    return ctx->fail(HARK_ERR_ARG, std::string("unknown option: ") + key);
}

extern "C" int hark_context_get_option(hark_ctx *ctx, const char *key, int64_t *value) {
    if (!ctx || !key || !value) return HARK_ERR_ARG;
    auto it = ctx->counters.find(key);
    if (it == ctx->counters.end()) {
        it = ctx->opts.find(key);
        if (it == ctx->opts.end()) return ctx->fail(HARK_ERR_ARG, std::string("option not set: ") + key);
    }
    *value = it->second;
    return HARK_OK;
}

extern "C" int hark_stats_last(hark_ctx *ctx, hark_stats *out) {
    HK_ENTER(ctx);
    if (!out) return ctx->fail(HARK_ERR_ARG, "stats: null output");
    if (ctx->stats_pending) {
        HK_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
        float ms = 0.f;
        HK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
        ctx->last.total_ms = ms;
        HK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1));
        ctx->last.kernel_ms = ms;
        ctx->stats_pending = false;
    }
    *out = ctx->last;
    return HARK_OK;
}

extern "C" int64_t hark_stats_total_launches(hark_ctx *ctx) { return ctx ? ctx->total_launches : -1; }

extern "C" void *hark_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (bytes <= 0) bytes = 1;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void hark_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------
int hk_table_alloc(hark_ctx *ctx, hark_table **out, int64_t n, int64_t cap, const int32_t *dtypes, int64_t m) {
    if (cap < n) cap = n;
    hark_table *t = new hark_table();
    t->n = n;
    t->cap = cap;
    t->cols.resize((size_t)m);
    for (int64_t c = 0; c < m; c++) {
        t->cols[c].dtype = dtypes[c];
        t->cols[c].owned = true;
        int rc = ctx->dalloc(&t->cols[c].ptr, (size_t)std::max<int64_t>(cap, 1) * hk_dtype_size(dtypes[c]));
        if (rc != HARK_OK) {
            for (int64_t j = 0; j < c; j++) ctx->dfree(t->cols[j].ptr);
            delete t;
            return rc;
        }
    }
    *out = t;
    return HARK_OK;
}

extern "C" int hark_table_free(hark_ctx *ctx, hark_table *t) {
    HK_ENTER(ctx);
    hk_table_free(ctx, t);
    return HARK_OK;
}

extern "C" int hark_table_shape(hark_ctx *ctx, const hark_table *t, int64_t shape[2]) {
    if (!ctx) return HARK_ERR_ARG;
    if (!t || !shape) return ctx->fail(HARK_ERR_ARG, "table_shape: null argument");
    shape[0] = t->n;
    shape[1] = (int64_t)t->cols.size();
    return HARK_OK;
}

extern "C" int hark_table_dtypes(hark_ctx *ctx, const hark_table *t, int32_t *dtypes_out) {
    if (!ctx) return HARK_ERR_ARG;
    if (!t || !dtypes_out) return ctx->fail(HARK_ERR_ARG, "table_dtypes: null argument");
    for (size_t c = 0; c < t->cols.size(); c++) dtypes_out[c] = t->cols[c].dtype;
    return HARK_OK;
}

extern "C" void *hark_table_column_ptr(hark_ctx *ctx, const hark_table *t, int32_t col) {
    if (!ctx || !t || col < 0 || (size_t)col >= t->cols.size()) return nullptr;
    return t->cols[col].ptr;
}

// ---- row-major <-> SoA transposition through shared memory ----
constexpr int TR_THREADS = 256;
constexpr int TR_MAX_COLS = 64;    // columns per launch
constexpr int TR_SMEM_ELEMS = 6144; // staged elements per block (48 KB at 8 bytes)

struct hk_colptrs {
    void *p[TR_MAX_COLS];
};

// rows [0, nrows) of a row-major chunk `rm` (row stride m_total elements), columns [c0, c0+mc)
// -> cols.p[j][dst_row0 + r].  Block b stages R rows; loads are contiguous when mc == m_total.
template <typename T>
__global__ void __launch_bounds__(TR_THREADS) hk_rm_to_soa_kernel(const T *__restrict__ rm, hk_colptrs cols,
                                                                   int64_t nrows, int m_total, int c0, int mc, int R,
                                                                   int64_t dst_row0) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    T *s = reinterpret_cast<T *>(tr_smem);
    const int stride = mc | 1;
    const int64_t r0 = (int64_t)blockIdx.x * R;
    const int rows_here = (int)min((int64_t)R, nrows - r0);
    const int elems = rows_here * mc;
    const T *src = rm + r0 * m_total + c0;
    for (int i = threadIdx.x; i < elems; i += TR_THREADS) {
        const int r = i / mc, c = i - r * mc;
        s[r * stride + c] = src[(int64_t)r * m_total + c];
    }
    __syncthreads();
    for (int c = 0; c < mc; c++) {
        T *dst = reinterpret_cast<T *>(cols.p[c]) + dst_row0 + r0;
        for (int r = threadIdx.x; r < rows_here; r += TR_THREADS) dst[r] = s[r * stride + c];
    }
}

template <typename T>
__global__ void __launch_bounds__(TR_THREADS) hk_soa_to_rm_kernel(T *__restrict__ rm, hk_colptrs cols, int64_t nrows,
                                                                   int m_total, int c0, int mc, int R,
                                                                   int64_t src_row0) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    T *s = reinterpret_cast<T *>(tr_smem);
    const int stride = mc | 1;
    const int64_t r0 = (int64_t)blockIdx.x * R;
    const int rows_here = (int)min((int64_t)R, nrows - r0);
    for (int c = 0; c < mc; c++) {
        const T *src = reinterpret_cast<const T *>(cols.p[c]) + src_row0 + r0;
        for (int r = threadIdx.x; r < rows_here; r += TR_THREADS) s[r * stride + c] = src[r];
    }
    __syncthreads();
    const int elems = rows_here * mc;
    T *dst = rm + r0 * m_total + c0;
    for (int i = threadIdx.x; i < elems; i += TR_THREADS) {
        const int r = i / mc, c = i - r * mc;
        dst[(int64_t)r * m_total + c] = s[r * stride + c];
    }
}

static int tr_rows_per_block(int mc) {
    int R = TR_SMEM_ELEMS / (mc | 1);
    R = std::min(R, 1024);
    R = std::max(32, R / 32 * 32);
    return R;
}

// Transposes `nrows` rows between a device row-major chunk and table columns, for a homogeneous table.
static int hk_transpose_chunk(hark_ctx *ctx, bool to_soa, void *rm, const hark_table *t, int64_t nrows,
                              int64_t tbl_row0) {
    const int m = (int)t->cols.size();
    const int w = hk_dtype_size(t->cols[0].dtype);
    if (nrows == 0 || m == 0) return HARK_OK;
    for (int c0 = 0; c0 < m; c0 += TR_MAX_COLS) {
        const int mc = std::min(TR_MAX_COLS, m - c0);
        hk_colptrs cp;
        for (int j = 0; j < mc; j++) cp.p[j] = t->cols[c0 + j].ptr;
        const int R = tr_rows_per_block(mc);
        const size_t smem = (size_t)R * (mc | 1) * w;
        const unsigned grid = (unsigned)((nrows + R - 1) / R);
        if (to_soa) {
            if (w == 4)
                hk_rm_to_soa_kernel<uint32_t><<<grid, TR_THREADS, smem, ctx->stream>>>((const uint32_t *)rm, cp, nrows, m, c0, mc, R, tbl_row0);
            else
                hk_rm_to_soa_kernel<uint64_t><<<grid, TR_THREADS, smem, ctx->stream>>>((const uint64_t *)rm, cp, nrows, m, c0, mc, R, tbl_row0);
        } else {
            if (w == 4)
                hk_soa_to_rm_kernel<uint32_t><<<grid, TR_THREADS, smem, ctx->stream>>>((uint32_t *)rm, cp, nrows, m, c0, mc, R, tbl_row0);
            else
                hk_soa_to_rm_kernel<uint64_t><<<grid, TR_THREADS, smem, ctx->stream>>>((uint64_t *)rm, cp, nrows, m, c0, mc, R, tbl_row0);
        }
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
    }
    return HARK_OK;
}

static bool table_homogeneous(const hark_table *t) {
    for (auto &c : t->cols)
        if (c.dtype != t->cols[0].dtype) return false;
    return true;
}

extern "C" int hark_table_from_host(hark_ctx *ctx, hark_table **out, const void *rowmajor, int64_t n, int64_t m,
                                    int32_t dtype) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out != nullptr, "table_from_host: null output");
    HK_ARG(ctx, n >= 0 && m >= 0 && m <= INT32_MAX, "table_from_host: bad shape");
    HK_ARG(ctx, hk_dtype_ok(dtype), "table_from_host: bad dtype");
    HK_ARG(ctx, rowmajor != nullptr || n * m == 0, "table_from_host: null data");
    std::vector<int32_t> dts((size_t)m, dtype);
    hark_table *t = nullptr;
    HK_TRY(hk_table_alloc(ctx, &t, n, n, dts.data(), m));
    if (n * m == 0) {
        *out = t;
        return HARK_OK;
    }
    // double-buffered: H2D of chunk i+1 (copy stream) overlaps the transpose of chunk i (ctx stream)
    const int w = hk_dtype_size(dtype);
    const size_t row_bytes = (size_t)m * w;
    const int64_t chunk_mb = ctx->opt("upload.chunk_mb", 64);
    int64_t chunk_rows = std::max<int64_t>(1, (chunk_mb << 20) / (int64_t)row_bytes);
    chunk_rows = std::min(chunk_rows, n);
    void *stage[2] = {nullptr, nullptr};
    const int nbuf = chunk_rows < n ? 2 : 1;
    int rc = HARK_OK;
    for (int b = 0; b < nbuf && rc == HARK_OK; b++) rc = ctx->dalloc(&stage[b], (size_t)chunk_rows * row_bytes);
    cudaError_t e = cudaSuccess;
    // the staging buffers were allocated in ctx->stream order; make the copy stream see them
    if (rc == HARK_OK) e = cudaEventRecord(ctx->ev_done[0], ctx->stream);
    if (rc == HARK_OK && e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[0], 0);
    int i = 0;
    for (int64_t r0 = 0; rc == HARK_OK && e == cudaSuccess && r0 < n; r0 += chunk_rows, i++) {
        const int b = i & 1;
        const int64_t rows = std::min(chunk_rows, n - r0);
        if (i >= 2) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(stage[b], (const char *)rowmajor + (size_t)r0 * row_bytes, (size_t)rows * row_bytes,
                                cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0);
        if (e == cudaSuccess) rc = hk_transpose_chunk(ctx, true, stage[b], t, rows, r0);
        if (e == cudaSuccess && rc == HARK_OK) e = cudaEventRecord(ctx->ev_done[b], ctx->stream);
    }
    for (int b = 0; b < nbuf; b++) ctx->dfree(stage[b]);
    // the caller may reuse `rowmajor` as soon as we return: drain the H2D copies (pinned memory
    // makes them truly asynchronous); the transposes may still be running on the device
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->copy_stream);
    if (e != cudaSuccess && rc == HARK_OK)
        rc = ctx->fail(HARK_ERR_CUDA, std::string("table_from_host: ") + cudaGetErrorString(e));
    if (rc != HARK_OK) {
        hk_table_free(ctx, t);
        return rc;
    }
    *out = t;
    return HARK_OK;
    HK_ABI_END(ctx)
}

extern "C" int hark_table_from_columns(hark_ctx *ctx, hark_table **out, const void *const *host_cols,
                                       const int32_t *dtypes, int64_t n, int64_t m) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && (m == 0 || (host_cols && dtypes)) && n >= 0 && m >= 0, "table_from_columns: bad argument");
    for (int64_t c = 0; c < m; c++) HK_ARG(ctx, hk_dtype_ok(dtypes[c]), "table_from_columns: bad dtype");
    hark_table *t = nullptr;
    HK_TRY(hk_table_alloc(ctx, &t, n, n, dtypes, m));
    for (int64_t c = 0; c < m && n > 0; c++) {
        cudaError_t e = cudaMemcpyAsync(t->cols[c].ptr, host_cols[c], (size_t)n * hk_dtype_size(dtypes[c]),
                                        cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
            hk_table_free(ctx, t);
            return ctx->fail(HARK_ERR_CUDA, std::string("table_from_columns: ") + cudaGetErrorString(e));
        }
    }
    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // host buffers may be reused on return
    *out = t;
    return HARK_OK;
    HK_ABI_END(ctx)
}

extern "C" int hark_table_from_device(hark_ctx *ctx, hark_table **out, void *const *dev_cols, const int32_t *dtypes,
                                      int64_t n, int64_t m) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && (m == 0 || (dev_cols && dtypes)) && n >= 0 && m >= 0, "table_from_device: bad argument");
    hark_table *t = new hark_table();
    t->n = n;
    t->cap = n;
    t->cols.resize((size_t)m);
    for (int64_t c = 0; c < m; c++) {
        if (!hk_dtype_ok(dtypes[c]) || ((uintptr_t)dev_cols[c] & 15) != 0 || (!dev_cols[c] && n > 0)) {
            delete t;
            return ctx->fail(HARK_ERR_ARG, "table_from_device: bad dtype or column pointer not 16-byte aligned");
        }
        t->cols[c].ptr = dev_cols[c];
        t->cols[c].dtype = dtypes[c];
        t->cols[c].owned = false;
    }
    *out = t;
    return HARK_OK;
    HK_ABI_END(ctx)
}

extern "C" int hark_table_to_host(hark_ctx *ctx, const hark_table *t, void *rowmajor_out) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, t != nullptr, "table_to_host: null table");
    const int64_t n = t->n, m = (int64_t)t->cols.size();
    if (n * m == 0) return HARK_OK;
    HK_ARG(ctx, rowmajor_out != nullptr, "table_to_host: null output");
    HK_ARG(ctx, table_homogeneous(t), "table_to_host: columns differ in dtype; use hark_table_column_to_host");
    const int w = hk_dtype_size(t->cols[0].dtype);
    const size_t row_bytes = (size_t)m * w;
    const int64_t chunk_mb = ctx->opt("upload.chunk_mb", 64);
    int64_t chunk_rows = std::min<int64_t>(n, std::max<int64_t>(1, (chunk_mb << 20) / (int64_t)row_bytes));
    void *stage[2] = {nullptr, nullptr};
    const int nbuf = chunk_rows < n ? 2 : 1;
    int rc = HARK_OK;
    for (int b = 0; b < nbuf && rc == HARK_OK; b++) rc = ctx->dalloc(&stage[b], (size_t)chunk_rows * row_bytes);
    cudaError_t e = cudaSuccess;
    int i = 0;
    for (int64_t r0 = 0; rc == HARK_OK && e == cudaSuccess && r0 < n; r0 += chunk_rows, i++) {
        const int b = i & 1;
        const int64_t rows = std::min(chunk_rows, n - r0);
        if (i >= 2) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0); // D2H of chunk i-2 drained
        if (e == cudaSuccess) rc = hk_transpose_chunk(ctx, false, stage[b], t, rows, r0);
        if (e == cudaSuccess && rc == HARK_OK) e = cudaEventRecord(ctx->ev_done[b], ctx->stream);
        if (e == cudaSuccess && rc == HARK_OK) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0);
        if (e == cudaSuccess && rc == HARK_OK)
            e = cudaMemcpyAsync((char *)rowmajor_out + (size_t)r0 * row_bytes, stage[b], (size_t)rows * row_bytes,
                                cudaMemcpyDeviceToHost, ctx->copy_stream);
        if (e == cudaSuccess && rc == HARK_OK) e = cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    for (int b = 0; b < nbuf; b++) ctx->dfree(stage[b]);
    if (e != cudaSuccess && rc == HARK_OK)
        rc = ctx->fail(HARK_ERR_CUDA, std::string("table_to_host: ") + cudaGetErrorString(e));
    return rc;
    HK_ABI_END(ctx)
}

extern "C" int hark_table_column_to_host(hark_ctx *ctx, const hark_table *t, int32_t col, int64_t row0, int64_t nrows,
                                         void *out) {
    HK_ENTER(ctx);
    HK_ARG(ctx, t && col >= 0 && (size_t)col < t->cols.size(), "column_to_host: bad column");
    HK_ARG(ctx, row0 >= 0 && nrows >= 0 && row0 + nrows <= t->n, "column_to_host: bad row range");
    if (nrows == 0) return HARK_OK;
    HK_ARG(ctx, out != nullptr, "column_to_host: null output");
    const int w = hk_dtype_size(t->cols[col].dtype);
    HK_CUDA(ctx, cudaMemcpyAsync(out, (const char *)t->cols[col].ptr + (size_t)row0 * w, (size_t)nrows * w,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HARK_OK;
}

// ------------------------------------------------------------------------------------------
// synthetic generator — bit-identical twin of oracle.c oracle_synth_column
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hk_synth_kernel(void *__restrict__ out, int dtype, hark_colspec spec,
                                                        uint64_t seed, int col, int64_t row0, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t r = (uint64_t)(row0 + i);
        uint64_t iv = 0;
        double fv = 0.0;
        float ffv = 0.0f;
        if (spec.kind == HARK_GEN_UNIFORM) {
            const uint64_t h = hk_mix64(seed, (uint64_t)col, r);
            if (dtype == HARK_F32)
                ffv = __fmaf_rn((float)(h >> 40) * 0x1p-24f, (float)(spec.fhi - spec.flo), (float)spec.flo);
            else if (dtype == HARK_F64)
                fv = __fma_rn((double)(h >> 11) * 0x1p-53, spec.fhi - spec.flo, spec.flo);
            else
                iv = (uint64_t)spec.lo + (spec.range ? __umul64hi(h, spec.range) : h);
        } else if (spec.kind == HARK_GEN_AFFINE) {
            uint64_t v = spec.a * r + spec.b;
            if (spec.range) v %= spec.range;
            iv = v;
            fv = (double)v;
            ffv = (float)v;
        } else if (spec.kind == HARK_GEN_AFFINE_UNIFORM) {
            const uint64_t v = spec.a * __umul64hi(hk_mix64(seed, (uint64_t)col, r), spec.range) + spec.b;
            iv = v;
            fv = (double)v;
            ffv = (float)v;
        } else if (spec.kind == HARK_GEN_LOGUNIFORM) {
            const uint64_t range = spec.range < 2 ? 2 : spec.range;
            const int nb = 63 - __clzll((long long)range); // floor(log2 range) >= 1
            const uint64_t e = __umul64hi(hk_mix64(seed, (uint64_t)col, r), (uint64_t)nb);
            const uint64_t h2 = hk_mix64(seed ^ 0x5851F42D4C957F2DULL, (uint64_t)col, r);
            const uint64_t k = ((1ull << e) + (h2 & ((1ull << e) - 1ull)) - 1ull) % range;
            iv = (uint64_t)spec.lo + k;
            fv = (double)(long long)iv;
            ffv = (float)(long long)iv;
        } else {
            iv = (uint64_t)spec.lo;
            fv = spec.flo;
            ffv = (float)spec.flo;
        }
        switch (dtype) {
        case HARK_I32:
        case HARK_U32: ((uint32_t *)out)[i] = (uint32_t)iv; break;
        case HARK_I64: ((uint64_t *)out)[i] = iv; break;
        case HARK_F32: ((float *)out)[i] = ffv; break;
        default: ((double *)out)[i] = fv; break;
        }
    }
}

extern "C" int hark_table_synth(hark_ctx *ctx, hark_table **out, int64_t n, int64_t m, const int32_t *dtypes,
                                uint64_t seed, const hark_colspec *specs, int64_t row0) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && n >= 0 && m >= 0 && (m == 0 || (dtypes && specs)), "table_synth: bad argument");
    for (int64_t c = 0; c < m; c++) {
        HK_ARG(ctx, hk_dtype_ok(dtypes[c]), "table_synth: bad dtype");
        HK_ARG(ctx, specs[c].kind >= HARK_GEN_UNIFORM && specs[c].kind <= HARK_GEN_AFFINE_UNIFORM, "table_synth: bad kind");
    }
    hark_table *t = nullptr;
    HK_TRY(hk_table_alloc(ctx, &t, n, n, dtypes, m));
    if (n > 0) {
        const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 16);
        for (int64_t c = 0; c < m; c++) {
            hk_synth_kernel<<<grid, 256, 0, ctx->stream>>>(t->cols[c].ptr, dtypes[c], specs[c], seed, (int)c, row0, n);
            ctx->count_launch();
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            hk_table_free(ctx, t);
            return ctx->fail(HARK_ERR_CUDA, std::string("table_synth: ") + cudaGetErrorString(e));
        }
    }
    *out = t;
    return HARK_OK;
    HK_ABI_END(ctx)
}

// ------------------------------------------------------------------------------------------
// slice / concat
// ------------------------------------------------------------------------------------------
extern "C" int hark_table_slice(hark_ctx *ctx, hark_table **out, const hark_table *t, int64_t row0, int64_t nrows) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && t && row0 >= 0 && nrows >= 0 && row0 + nrows <= t->n, "table_slice: bad argument");
    const int64_t m = (int64_t)t->cols.size();
    std::vector<int32_t> dts, idx;
    for (int64_t c = 0; c < m; c++) {
        dts.push_back(t->cols[c].dtype);
        idx.push_back((int32_t)c);
    }
    hark_table *r = nullptr;
    HK_TRY(hk_table_alloc(ctx, &r, nrows, nrows, dts.data(), m));
    int rc = hk_copy_columns(ctx, r, t, idx.data(), m, row0, nrows, 0);
    if (rc != HARK_OK) {
        hk_table_free(ctx, r);
        return rc;
    }
    *out = r;
    return HARK_OK;
    HK_ABI_END(ctx)
}

extern "C" int hark_table_concat(hark_ctx *ctx, hark_table **out, const hark_table *a, const hark_table *b) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && a && b && a->cols.size() == b->cols.size(), "table_concat: schemas differ");
    const int64_t m = (int64_t)a->cols.size();
    std::vector<int32_t> dts, idx;
    for (int64_t c = 0; c < m; c++) {
        HK_ARG(ctx, a->cols[c].dtype == b->cols[c].dtype, "table_concat: schemas differ");
        dts.push_back(a->cols[c].dtype);
        idx.push_back((int32_t)c);
    }
    hark_table *r = nullptr;
    HK_TRY(hk_table_alloc(ctx, &r, a->n + b->n, a->n + b->n, dts.data(), m));
    int rc = hk_copy_columns(ctx, r, a, idx.data(), m, 0, a->n, 0);
    if (rc == HARK_OK) rc = hk_copy_columns(ctx, r, b, idx.data(), m, 0, b->n, a->n);
    if (rc != HARK_OK) {
        hk_table_free(ctx, r);
        return rc;
    }
    *out = r;
    return HARK_OK;
    HK_ABI_END(ctx)
}
