// repartition.cu — the device half of the multi-GPU exchange step:
//   hark_table_partition_by_splitters  K8b: stable range partition of a shard into per-destination contiguous regions
//   hark_table_sample_order_keys       order keys of sampled rows (splitter candidates)
//   hark_entry_groupby_finalize        AVG = merged f64 sum / merged count after the partial-aggregate merge
// The reference is single-process and has no counterpart; ordering semantics are those of the
// single-GPU operators: splitters compare through the same order-key mapping as ORDER BY (orderby.cu).
#include <algorithm>
#include <new>
#include <stdexcept>
#include <vector>

#include "hark_internal.cuh"
#include "sort.cuh"

namespace {

constexpr int RMAXK = 4;      // key columns in a splitter tuple
constexpr int RMAXP = 256;    // destinations

struct KeyCols {
    int nk;
    const void *ptr[RMAXK];
    int dtype[RMAXK];
    int desc[RMAXK];
};

__device__ __forceinline__ uint64_t row_ordkey(const KeyCols &K, int j, int64_t r) {
    uint64_t u;
    const int dt = K.dtype[j];
    if (dt == HARK_I64 || dt == HARK_F64) u = hk_ordkey64(reinterpret_cast<const unsigned long long *>(K.ptr[j])[r], dt);
    else u = (uint64_t)hk_ordkey32(reinterpret_cast<const uint32_t *>(K.ptr[j])[r], dt);
    return K.desc[j] ? ~u : u;
}

// digit[r] = number of splitter tuples <= key tuple of row r (upper bound), counts[digit]++
__global__ void __launch_bounds__(256) hk_splitter_digit_kernel(KeyCols K, int64_t n, const unsigned long long *__restrict__ splitters,
                                                                 int nsplit, uint32_t *__restrict__ digit,
                                                                 unsigned long long *__restrict__ counts) {
    __shared__ unsigned long long s_sp[(RMAXP - 1) * RMAXK];
    __shared__ uint32_t s_cnt[RMAXP];
    for (int i = threadIdx.x; i < nsplit * K.nk; i += 256) s_sp[i] = splitters[i];
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        uint64_t key[RMAXK];
#pragma unroll
        for (int j = 0; j < RMAXK; j++) key[j] = j < K.nk ? row_ordkey(K, j, r) : 0ull;
        int lo = 0, hi = nsplit; // first splitter > key
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            bool le = true; // splitter[mid] <= key ?
#pragma unroll
            for (int j = 0; j < RMAXK; j++) {
                if (j < K.nk) {
                    const unsigned long long s = s_sp[mid * K.nk + j];
                    if (s != key[j]) {
                        le = s < key[j];
                        break;
                    }
                }
            }
            if (le) lo = mid + 1; else hi = mid;
        }
        digit[r] = (uint32_t)lo;
        atomicAdd(&s_cnt[lo], 1u);
    }
    __syncthreads();
    if (threadIdx.x <= nsplit && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(256) hk_sample_keys_kernel(KeyCols K, const long long *__restrict__ rows, int64_t nrows,
                                                              unsigned long long *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    for (int j = 0; j < K.nk; j++) out[i * K.nk + j] = row_ordkey(K, j, rows[i]);
}

__global__ void __launch_bounds__(256) hk_divide_kernel(double *__restrict__ out, const double *__restrict__ sum,
                                                         const long long *__restrict__ cnt, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = sum[i] / (double)cnt[i];
}

unsigned grid_for(hark_ctx *ctx, int64_t n, int per_sm = 8) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * per_sm));
}

int fill_keycols(hark_ctx *ctx, KeyCols &K, const hark_table *db, const int32_t *key_cols, const int32_t *desc, int64_t nk) {
    const int64_t m = (int64_t)db->cols.size();
    HK_ARG(ctx, nk >= 1 && nk <= RMAXK, "repartition: 1..4 key columns");
    memset(&K, 0, sizeof K);
    K.nk = (int)nk;
    for (int64_t j = 0; j < nk; j++) {
        HK_ARG(ctx, key_cols[j] >= 0 && key_cols[j] < m, "repartition: key column index out of bounds");
        K.ptr[j] = db->cols[key_cols[j]].ptr;
        K.dtype[j] = db->cols[key_cols[j]].dtype;
        K.desc[j] = desc ? (desc[j] != 0) : 0;
    }
    return HARK_OK;
}

} // namespace

#define HK_ENTER(ctx)                \
    if (!(ctx)) return HARK_ERR_ARG; \
    (ctx)->entry_depth = 0;          \
    HK_CUDA(ctx, cudaSetDevice((ctx)->device))

extern "C" int hark_table_partition_by_splitters(hark_ctx *ctx, hark_table **out, const hark_table *db,
                                                 const int32_t *key_cols, const int32_t *desc, int64_t nk,
                                                 const uint64_t *splitters, int32_t nparts, int64_t *counts_out) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && db && key_cols && counts_out && nparts >= 1 && nparts <= RMAXP && (nparts == 1 || splitters),
           "partition_by_splitters: bad argument");
    const int64_t n = db->n, m = (int64_t)db->cols.size();
    KeyCols K;
    HK_TRY(fill_keycols(ctx, K, db, key_cols, desc, nk));
    ctx->entry_begin();
    const int nsplit = nparts - 1;
    unsigned long long *d_sp = nullptr, *d_cnt = nullptr;
    uint32_t *d_digit = nullptr;
    struct Tmp {
        hark_ctx *ctx;
        std::vector<void *> v;
        ~Tmp() {
            for (void *p : v) ctx->dfree(p);
        }
    } tmp{ctx, {}};
    HK_TRY(ctx->dalloc((void **)&d_sp, sizeof(unsigned long long) * (size_t)std::max<int64_t>(1, (int64_t)nsplit * nk)));
    tmp.v.push_back(d_sp);
    HK_TRY(ctx->dalloc((void **)&d_cnt, sizeof(unsigned long long) * RMAXP));
    tmp.v.push_back(d_cnt);
    HK_TRY(ctx->dalloc((void **)&d_digit, sizeof(uint32_t) * (size_t)std::max<int64_t>(n, 1)));
    tmp.v.push_back(d_digit);
    if (nsplit > 0)
        HK_CUDA(ctx, cudaMemcpyAsync(d_sp, splitters, sizeof(uint64_t) * (size_t)nsplit * nk, cudaMemcpyHostToDevice, ctx->stream));
    HK_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * RMAXP, ctx->stream));
    if (n > 0) {
        hk_splitter_digit_kernel<<<grid_for(ctx, n, 8), 256, 0, ctx->stream>>>(K, n, d_sp, nsplit, d_digit, d_cnt);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
    }
    HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, d_cnt, sizeof(uint64_t) * (size_t)nparts, cudaMemcpyDeviceToHost, ctx->stream));
    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // also: the host splitter buffer is free to go
    for (int p = 0; p < nparts; p++) counts_out[p] = (int64_t)ctx->h_scalars[p];

    // one stable radix pass keyed on the destination digit; every column rides along
    std::vector<hk_sort_array> arrays;
    {
        hk_sort_array a;
        a.in = d_digit;
        a.width = 4;
        arrays.push_back(a);
    }
    const bool direct = m + 1 <= HK_SORT_MAX_ARRAYS;
    void *rowid_in = nullptr;
    const int rw = n > 0xffffffffll ? 8 : 4;
    if (direct) {
        for (int64_t c = 0; c < m; c++) {
            hk_sort_array a;
            a.in = db->cols[c].ptr;
            a.width = hk_dtype_size(db->cols[c].dtype);
            arrays.push_back(a);
        }
    } else {
        HK_TRY(ctx->dalloc(&rowid_in, (size_t)std::max<int64_t>(n, 1) * rw));
        tmp.v.push_back(rowid_in);
        HK_TRY(hk_iota(ctx, rowid_in, n, rw));
        hk_sort_array r;
        r.in = rowid_in;
        r.width = rw;
        arrays.push_back(r);
    }
    std::vector<hk_sort_keyspec> keys{hk_sort_keyspec{0, HARK_U32, 0}};
    ctx->kernel_begin();
    HK_TRY(hk_radix_sort(ctx, n, keys, arrays, 0, nullptr, nullptr));
    ctx->kernel_end();
    ctx->dfree(arrays[0].result);
    hark_table *t = new hark_table();
    t->n = n;
    t->cap = n;
    int rc = HARK_OK;
    if (direct) {
        for (int64_t c = 0; c < m; c++) {
            hark_col col;
            col.ptr = arrays[1 + c].result;
            col.dtype = db->cols[c].dtype;
            col.owned = true;
            t->cols.push_back(col);
        }
    } else {
        for (int64_t c = 0; c < m && rc == HARK_OK; c++) {
            const int w = hk_dtype_size(db->cols[c].dtype);
            void *p = nullptr;
            rc = ctx->dalloc(&p, (size_t)std::max<int64_t>(n, 1) * w);
            if (rc != HARK_OK) break;
            hark_col col;
            col.ptr = p;
            col.dtype = db->cols[c].dtype;
            col.owned = true;
            t->cols.push_back(col);
            rc = hk_gather(ctx, p, db->cols[c].ptr, w, arrays[1].result, rw, n);
        }
        ctx->dfree(arrays[1].result);
        if (rc != HARK_OK) {
            hk_table_free(ctx, t);
            return rc;
        }
    }
    int64_t alg = 0;
    for (int64_t c = 0; c < m; c++) alg += 2 * n * hk_dtype_size(db->cols[c].dtype);
    ctx->entry_end(alg, n, n);
    *out = t;
    return HARK_OK;
    HK_ABI_END(ctx)
}

extern "C" int hark_table_sample_order_keys(hark_ctx *ctx, const hark_table *db, const int32_t *key_cols,
                                            const int32_t *desc, int64_t nk, const int64_t *rows, int64_t nrows,
                                            uint64_t *out) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, db && key_cols && nrows >= 0 && (nrows == 0 || (rows && out)), "sample_order_keys: bad argument");
    KeyCols K;
    HK_TRY(fill_keycols(ctx, K, db, key_cols, desc, nk));
    if (nrows == 0) return HARK_OK;
    for (int64_t i = 0; i < nrows; i++) HK_ARG(ctx, rows[i] >= 0 && rows[i] < db->n, "sample_order_keys: row out of range");
    long long *d_rows = nullptr;
    unsigned long long *d_out = nullptr;
    HK_TRY(ctx->dalloc((void **)&d_rows, sizeof(long long) * (size_t)nrows));
    int rc = ctx->dalloc((void **)&d_out, sizeof(unsigned long long) * (size_t)(nrows * nk));
    if (rc != HARK_OK) {
        ctx->dfree(d_rows);
        return rc;
    }
    cudaError_t e = cudaMemcpyAsync(d_rows, rows, sizeof(long long) * (size_t)nrows, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        hk_sample_keys_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, ctx->stream>>>(K, d_rows, nrows, d_out);
        e = cudaGetLastError();
        ctx->count_launch();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(uint64_t) * (size_t)(nrows * nk), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ctx->dfree(d_rows);
    ctx->dfree(d_out);
    if (e != cudaSuccess) return ctx->fail(HARK_ERR_CUDA, std::string("sample_order_keys: ") + cudaGetErrorString(e));
    return HARK_OK;
    HK_ABI_END(ctx)
}

extern "C" int hark_entry_groupby_finalize(hark_ctx *ctx, hark_table **out, const hark_table *merged, const int32_t *ops,
                                           int64_t c) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, out && merged && c >= 0 && (c == 0 || ops), "groupby_finalize: bad argument");
    const int64_t m = (int64_t)merged->cols.size(), G = merged->n;
    int64_t need = 1;
    for (int64_t j = 0; j < c; j++) need += ops[j] == HARK_AGG_AVG ? 2 : 1;
    HK_ARG(ctx, m == need, "groupby_finalize: partial layout does not match the aggregate list");
    std::vector<int32_t> odt{merged->cols[0].dtype};
    {
        int64_t col = 1;
        for (int64_t j = 0; j < c; j++) {
            if (ops[j] == HARK_AGG_AVG) {
                HK_ARG(ctx, merged->cols[col].dtype == HARK_F64 && merged->cols[col + 1].dtype == HARK_I64,
                       "groupby_finalize: AVG needs an f64 sum and an i64 count");
                odt.push_back(HARK_F64);
                col += 2;
            } else {
                odt.push_back(merged->cols[col].dtype);
                col += 1;
            }
        }
    }
    ctx->entry_begin();
    hark_table *t = nullptr;
    HK_TRY(hk_table_alloc(ctx, &t, G, G, odt.data(), 1 + c));
    int rc = HARK_OK;
    int64_t col = 0;
    for (int64_t j = -1; j < c && rc == HARK_OK; j++) {
        hark_table dst_view; // single-column borrowed view so that hk_copy_columns writes output column 1+j
        dst_view.n = G;
        dst_view.cap = G;
        hark_col dc = t->cols[(size_t)(j + 1)];
        dc.owned = false;
        dst_view.cols.push_back(dc);
        if (j >= 0 && ops[j] == HARK_AGG_AVG) {
            if (G > 0) {
                hk_divide_kernel<<<grid_for(ctx, G), 256, 0, ctx->stream>>>((double *)dc.ptr, (const double *)merged->cols[col].ptr,
                                                                          (const long long *)merged->cols[col + 1].ptr, G);
                if (cudaGetLastError() != cudaSuccess) rc = ctx->fail(HARK_ERR_CUDA, "groupby_finalize: launch failed");
                ctx->count_launch();
            }
            col += 2;
        } else {
            const int32_t one = (int32_t)col;
            rc = hk_copy_columns(ctx, &dst_view, merged, &one, 1, 0, G, 0);
            col += 1;
        }
    }
    if (rc != HARK_OK) {
        hk_table_free(ctx, t);
        return rc;
    }
    ctx->entry_end(0, G, G);
    *out = t;
    return HARK_OK;
    HK_ABI_END(ctx)
}

// ------------------------------------------------------------------------------------------------------------
// K8c — partition fused with the exchange: every rank owns a receive ARENA (cudaMalloc, exported through CUDA IPC,
// mapped by the other ranks of the NVSwitch domain); the stable scatter pass writes each row straight into its
// final slot in the destination GPU's arena (NVLink peer stores, one contiguous run per destination and tile), so
// the repartition needs no staging copy and no NCCL transfer: read the shard once, store over NVLink once.
//   phase 1  hark_peer_scatter_count   destination digit of every row + rows per destination
//   (host)   all ranks all-gather their count vectors -> counts[src][dst]; every rank derives every arena layout
//   phase 2  hark_peer_scatter_run     chunk histogram + scan + scatter with per-destination base addresses
//   (host)   one stream-ordered collective = "all stores into my arena have landed"
//   phase 3  hark_peer_scatter_result  borrowed table over this rank's arena
// Row order at the destination: source rank, then source row order (the scatter is stable) — the same order the
// NCCL all-to-all path delivers, so ORDER BY stays stable across GPUs.
// ------------------------------------------------------------------------------------------------------------
struct hk_peer_state {
    void *arena = nullptr;           // this rank's receive arena
    int64_t arena_bytes = 0;
    int world = 0, rank = -1;
    void *base[HK_PEER_MAX] = {};    // every rank's arena as mapped here (base[rank] == arena)
    // between phases
    uint32_t *d_digit = nullptr;
    int64_t n = 0;
    std::vector<int32_t> dtypes;
    std::vector<int64_t> my_col_off; // byte offset of every column in MY arena for the current exchange
    int64_t my_rows = 0;
};

void hk_peer_destroy(hark_ctx *ctx) {
    hk_peer_state *p = ctx->peer;
    if (!p) return;
    for (int r = 0; r < p->world; r++)
        if (r != p->rank && p->base[r]) cudaIpcCloseMemHandle(p->base[r]);
    if (p->d_digit) ctx->dfree(p->d_digit);
    if (p->arena) cudaFree(p->arena);
    delete p;
    ctx->peer = nullptr;
}

extern "C" int hark_peer_arena_create(hark_ctx *ctx, int64_t bytes, void *ipc_handle_out) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, bytes > 0 && ipc_handle_out, "peer_arena_create: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    hk_peer_destroy(ctx);
    hk_peer_state *p = new hk_peer_state();
    ctx->peer = p;
    bytes = (bytes + 4095) & ~(int64_t)4095;
    cudaError_t e = cudaMalloc(&p->arena, (size_t)bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        hk_peer_destroy(ctx);
        return ctx->fail(HARK_ERR_OOM, std::string("peer_arena_create: ") + cudaGetErrorString(e));
    }
    p->arena_bytes = bytes;
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p->arena);
    if (e != cudaSuccess) {
        cudaGetLastError();
        hk_peer_destroy(ctx);
        return ctx->fail(HARK_ERR_CUDA, std::string("peer_arena_create(ipc): ") + cudaGetErrorString(e));
    }
    memcpy(ipc_handle_out, &h, 64);
    return HARK_OK;
    HK_ABI_END(ctx)
}

extern "C" int hark_peer_arena_open(hark_ctx *ctx, const void *handles, int32_t world, int32_t my_rank) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, ctx->peer && ctx->peer->arena && handles && world >= 1 && world <= HK_PEER_MAX && my_rank >= 0 && my_rank < world,
           "peer_arena_open: bad argument (create the arena first; at most 16 ranks)");
    hk_peer_state *p = ctx->peer;
    p->world = world;
    p->rank = my_rank;
    for (int r = 0; r < world; r++) {
        if (r == my_rank) {
            p->base[r] = p->arena;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * r, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&p->base[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            p->base[r] = nullptr;
            return ctx->fail(HARK_ERR_CUDA, std::string("peer_arena_open: rank ") + std::to_string(r) + ": " + cudaGetErrorString(e));
        }
    }
    return HARK_OK;
    HK_ABI_END(ctx)
}

extern "C" int hark_peer_arena_close(hark_ctx *ctx) {
    HK_ENTER(ctx);
    hk_peer_destroy(ctx);
    return HARK_OK;
}

extern "C" int hark_peer_scatter_count(hark_ctx *ctx, const hark_table *db, const int32_t *key_cols, const int32_t *desc,
                                       int64_t nk, const uint64_t *splitters, int64_t *counts_out) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, ctx->peer && ctx->peer->world >= 1 && db && key_cols && counts_out, "peer_scatter_count: open the arenas first");
    hk_peer_state *p = ctx->peer;
    const int world = p->world, nsplit = world - 1;
    HK_ARG(ctx, nsplit == 0 || splitters, "peer_scatter_count: splitters missing");
    const int64_t n = db->n;
    KeyCols K;
    HK_TRY(fill_keycols(ctx, K, db, key_cols, desc, nk));
    ctx->entry_begin();
    if (p->d_digit) {
        ctx->dfree(p->d_digit);
        p->d_digit = nullptr;
    }
    unsigned long long *d_sp = nullptr, *d_cnt = nullptr;
    HK_TRY(ctx->dalloc((void **)&d_sp, sizeof(unsigned long long) * (size_t)std::max<int64_t>(1, (int64_t)nsplit * nk)));
    int rc = ctx->dalloc((void **)&d_cnt, sizeof(unsigned long long) * RMAXP);
    if (rc == HARK_OK) rc = ctx->dalloc((void **)&p->d_digit, sizeof(uint32_t) * (size_t)std::max<int64_t>(n, 1));
    cudaError_t e = cudaSuccess;
    if (rc == HARK_OK) {
        if (nsplit > 0) e = cudaMemcpyAsync(d_sp, splitters, sizeof(uint64_t) * (size_t)nsplit * nk, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * RMAXP, ctx->stream);
        if (e == cudaSuccess && n > 0) {
            hk_splitter_digit_kernel<<<grid_for(ctx, n, 8), 256, 0, ctx->stream>>>(K, n, d_sp, nsplit, p->d_digit, d_cnt);
            e = cudaGetLastError();
            ctx->count_launch();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_scalars, d_cnt, sizeof(uint64_t) * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    ctx->dfree(d_sp);
    ctx->dfree(d_cnt);
    if (rc != HARK_OK) return rc;
    if (e != cudaSuccess) return ctx->fail(HARK_ERR_CUDA, std::string("peer_scatter_count: ") + cudaGetErrorString(e));
    for (int d = 0; d < world; d++) counts_out[d] = (int64_t)ctx->h_scalars[d];
    p->n = n;
    p->dtypes.clear();
    for (auto &c : db->cols) p->dtypes.push_back(c.dtype);
    ctx->entry_end(0, n, n);
    return HARK_OK;
    HK_ABI_END(ctx)
}

// bytes of arena a destination needs for `rows` rows of this schema (every column 256-byte aligned)
static int64_t peer_layout(const std::vector<int32_t> &dtypes, int64_t rows, std::vector<int64_t> *col_off) {
    int64_t off = 0;
    if (col_off) col_off->clear();
    for (int32_t dt : dtypes) {
        if (col_off) col_off->push_back(off);
        off += (rows * hk_dtype_size(dt) + 255) & ~(int64_t)255;
        off += 256; // slack: 16-byte vector loads of a ragged column end stay inside the column's region
    }
    return off;
}

extern "C" int64_t hark_peer_arena_bytes_needed(hark_ctx *ctx, const hark_table *db, int64_t rows) {
    if (!ctx || !db) return -1;
    std::vector<int32_t> dts;
    for (auto &c : db->cols) dts.push_back(c.dtype);
    return peer_layout(dts, rows, nullptr);
}

extern "C" int hark_peer_scatter_run(hark_ctx *ctx, const hark_table *db, const int64_t *counts_matrix) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, ctx->peer && ctx->peer->d_digit && db && counts_matrix, "peer_scatter_run: call peer_scatter_count first");
    hk_peer_state *p = ctx->peer;
    const int world = p->world, me = p->rank;
    const int64_t n = db->n, m = (int64_t)db->cols.size();
    HK_ARG(ctx, n == p->n && m == (int64_t)p->dtypes.size(), "peer_scatter_run: table changed between the phases");
    HK_ARG(ctx, m + 1 <= HK_SORT_MAX_ARRAYS, "peer_scatter_run: too many columns");
    ctx->entry_begin();
    // layout of every destination's arena, and where MY rows start inside each column there
    std::vector<unsigned long long> h_out((size_t)(1 + m) * HK_PEER_MAX, 0ull);
    int64_t local_off = 0; // rows of mine bound for destinations < d (= position of bin d in my local order)
    for (int d = 0; d < world; d++) {
        int64_t rows_d = 0, before_me = 0;
        for (int s = 0; s < world; s++) {
            const int64_t c = counts_matrix[(size_t)s * world + d];
            HK_ARG(ctx, c >= 0, "peer_scatter_run: negative count");
            rows_d += c;
            if (s < me) before_me += c;
        }
        std::vector<int64_t> col_off;
        const int64_t need = peer_layout(p->dtypes, rows_d, &col_off);
        if (need > p->arena_bytes)
            return ctx->fail(HARK_ERR_OOM, "peer_scatter_run: a destination's rows do not fit its arena");
        for (int64_t c = 0; c < m; c++) {
            const int w = hk_dtype_size(p->dtypes[(size_t)c]);
            // element index used by the kernel = (position in my local order) ; slot wanted = before_me + (that - local_off)
            h_out[(size_t)(1 + c) * HK_PEER_MAX + d] =
                (unsigned long long)(uintptr_t)p->base[d] + (unsigned long long)col_off[(size_t)c] +
                (unsigned long long)((before_me - local_off) * (int64_t)w); // wraps; undone by the kernel's + position
        }
        if (d == me) {
            p->my_col_off = col_off;
            p->my_rows = rows_d;
        }
        local_off += counts_matrix[(size_t)me * world + d];
    }
    HK_ARG(ctx, local_off == n, "peer_scatter_run: counts do not add up to the shard's rows");
    unsigned long long *d_out = nullptr;
    HK_TRY(ctx->dalloc((void **)&d_out, sizeof(unsigned long long) * h_out.size()));
    // pinned staging not needed: the vector outlives the synchronous part of cudaMemcpyAsync from pageable memory
    cudaError_t e = cudaMemcpyAsync(d_out, h_out.data(), sizeof(unsigned long long) * h_out.size(), cudaMemcpyHostToDevice, ctx->stream);
    int rc = HARK_OK;
    if (e == cudaSuccess) {
        std::vector<const void *> cols;
        std::vector<int> widths;
        for (auto &c : db->cols) {
            cols.push_back(c.ptr);
            widths.push_back(hk_dtype_size(c.dtype));
        }
        ctx->kernel_begin();
        rc = hk_peer_scatter_pass(ctx, n, p->d_digit, cols.data(), widths.data(), (int)m, d_out);
        ctx->kernel_end();
    }
    ctx->dfree(d_out);
    ctx->dfree(p->d_digit);
    p->d_digit = nullptr;
    if (e != cudaSuccess) return ctx->fail(HARK_ERR_CUDA, std::string("peer_scatter_run: ") + cudaGetErrorString(e));
    if (rc != HARK_OK) return rc;
    int64_t alg = 0;
    for (auto &c : db->cols) alg += 2 * n * hk_dtype_size(c.dtype);
    ctx->entry_end(alg, n, n);
    return HARK_OK;
    HK_ABI_END(ctx)
}

extern "C" int hark_peer_scatter_result(hark_ctx *ctx, hark_table **out) {
    HK_ENTER(ctx);
    HK_ABI_BEGIN
    HK_ARG(ctx, ctx->peer && out && ctx->peer->my_col_off.size() == ctx->peer->dtypes.size() && !ctx->peer->dtypes.empty(),
           "peer_scatter_result: no exchange in flight");
    hk_peer_state *p = ctx->peer;
    hark_table *t = new hark_table();
    t->n = p->my_rows;
    t->cap = p->my_rows;
    for (size_t c = 0; c < p->dtypes.size(); c++) {
        hark_col col;
        col.ptr = (char *)p->arena + p->my_col_off[c];
        col.dtype = p->dtypes[c];
        col.owned = false; // a view of the arena: valid until the next exchange
        t->cols.push_back(col);
    }
    *out = t;
    return HARK_OK;
    HK_ABI_END(ctx)
}
