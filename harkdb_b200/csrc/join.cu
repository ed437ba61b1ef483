// join.cu — inner equi-join (reference semantics) and join + GROUP BY.
//
// Reference: futhark/join.fut:52-75 (an entry main.fut never imports): tag and concatenate both key columns
// (:55-57), stable 32-pass radix sort (:58), per-key segments (:59-64), a SEQUENTIAL loop that concatenates each
// key's cross product (:67-68), gather of the projected columns (:69-75).  Output order: key ascending
// (unsigned), then left row id, then right row id.
//
// Here: both sides are radix-sorted as (key, row id) pairs (stable, so equal keys keep row order); every sorted
// left row finds its match range in the sorted right keys by binary search (K6 count phase); an exclusive scan
// gives output offsets; the expand kernel maps every output row back to its (left, right) pair with a binary
// search over the offsets and gathers the projected columns straight from the SoA tables (K6 write phase).
// That reproduces the reference order exactly, without a sequential loop and for any duplication pattern.
#include <limits.h>

#include <algorithm>
#include <new>
#include <stdexcept>
#include <vector>

#include "dense_agg.cuh"
#include "hark_internal.cuh"
#include "sort.cuh"

int hk_segmented_aggregate(hark_ctx *ctx, hark_table **out, int64_t n, const void *sorted_key, int32_t key_dtype,
                           const std::vector<const void *> &vals, const std::vector<int32_t> &val_dtypes,
                           const std::vector<std::pair<int, int>> &aggs, bool pinned_u32);

namespace {

constexpr int JMAXC = 16;

// match range of every sorted left key in the sorted right keys
__global__ void __launch_bounds__(256) hk_join_bounds_kernel(const uint32_t *__restrict__ k1, int64_t n1,
                                                              const uint32_t *__restrict__ k2, int64_t n2,
                                                              uint32_t *__restrict__ lb_out, uint32_t *__restrict__ cnt_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += stride) {
        const uint32_t key = k1[i];
        int64_t lo = 0, hi = n2; // first index with k2 >= key
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (k2[mid] < key) lo = mid + 1; else hi = mid;
        }
        const int64_t lb = lo;
        hi = n2; // first index with k2 > key
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (k2[mid] <= key) lo = mid + 1; else hi = mid;
        }
        lb_out[i] = (uint32_t)lb;
        cnt_out[i] = (uint32_t)(lo - lb);
    }
}

// exclusive scan of u32 counts -> u64 offsets, total in offs[count]; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) hk_join_scan_kernel(const uint32_t *counts, unsigned long long *offs, int64_t count) {
    __shared__ unsigned long long s_part[1024];
    const int t = threadIdx.x;
    const int64_t per = (count + 1023) / 1024;
    const int64_t b = (int64_t)t * per, e = min(count, b + per);
    unsigned long long sum = 0;
    for (int64_t i = b; i < e; i++) sum += counts[i];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; i++) {
            const unsigned long long v = s_part[i];
            s_part[i] = run;
            run += v;
        }
        offs[count] = run;
    }
    __syncthreads();
    unsigned long long run = s_part[t];
    for (int64_t i = b; i < e; i++) {
        offs[i] = run;
        run += counts[i];
    }
}

struct ExpandParams {
    int64_t P, n1;
    const unsigned long long *offs; // [n1+1]
    const uint32_t *lb;             // [n1]
    const uint32_t *rid1, *rid2;    // sorted row ids
    int l, k;
    const uint32_t *src1[JMAXC];
    const uint32_t *src2[JMAXC];
    uint32_t *dst[2 * JMAXC];
};

__global__ void __launch_bounds__(256) hk_join_expand_kernel(const __grid_constant__ ExpandParams E) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < E.P; p += stride) {
        int64_t lo = 0, hi = E.n1; // last i with offs[i] <= p  (offs is non-decreasing, offs[n1] = P > p)
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (E.offs[mid] <= (unsigned long long)p) lo = mid; else hi = mid;
        }
        const int64_t i = lo;
        const int64_t j = p - (int64_t)E.offs[i];
        const uint32_t r1 = E.rid1[i];
        const uint32_t r2 = E.rid2[(int64_t)E.lb[i] + j];
        for (int c = 0; c < E.l; c++) E.dst[c][p] = E.src1[c][r1];
        for (int c = 0; c < E.k; c++) E.dst[E.l + c][p] = E.src2[c][r2];
    }
}

// ---- join + group by: dimension lookup ----
__device__ __forceinline__ long long load_int(const void *col, int dtype, int64_t i) {
    switch (dtype) {
    case HARK_I32: return (long long)reinterpret_cast<const int32_t *>(col)[i];
    case HARK_U32: return (long long)reinterpret_cast<const uint32_t *>(col)[i];
    default: return reinterpret_cast<const long long *>(col)[i];
    }
}

// lut[pk - pk_min] = dim row + 1 (0 = no such key).  Duplicate pks keep the largest row (flagged separately).
__global__ void __launch_bounds__(256) hk_lut_build_kernel(const void *pk, int pk_dtype, int64_t n_dim, long long pk_min,
                                                            uint32_t *lut, unsigned int *dup_flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_dim; i += stride) {
        const long long v = load_int(pk, pk_dtype, i) - pk_min;
        const uint32_t old = atomicMax(&lut[v], (uint32_t)(i + 1));
        if (old != 0) *dup_flag = 1u;
    }
}

// slot lookup for the fused probe + aggregate (dense_agg.cu, lut mode):
// lut[pk - pk_min] = (ordkey(dim.g) - g_lo) + 1, 0 = no such key.  Plain scattered stores (no atomics: an atomic's
// round trip to a lookup that does not fit L2 is what bounds this kernel); a duplicate pk then simply overwrites,
// and is caught afterwards because the number of non-empty entries falls short of the dimension's row count.
__global__ void __launch_bounds__(256) hk_slot_lut_build_kernel(const void *pk, int pk_dtype, const void *g, int g_dtype,
                                                                 int64_t n_dim, long long pk_min, unsigned long long g_lo,
                                                                 uint32_t *lut) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_dim; i += stride) {
        const long long v = load_int(pk, pk_dtype, i) - pk_min;
        unsigned long long ord;
        if (g_dtype == HARK_I64) ord = hk_ordkey64(reinterpret_cast<const unsigned long long *>(g)[i], g_dtype);
        else ord = hk_ordkey32(reinterpret_cast<const uint32_t *>(g)[i], g_dtype);
        lut[v] = (uint32_t)(ord - g_lo) + 1u;
    }
}

__global__ void __launch_bounds__(256) hk_count_nonzero_kernel(const uint32_t *__restrict__ a, int64_t n, unsigned long long *out) {
    unsigned long long c = 0;
    const int64_t nvec = n / 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 v = reinterpret_cast<const uint4 *>(a)[i];
        c += (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - nvec * 4)) c += a[nvec * 4 + threadIdx.x] != 0;
    c = hk_warp_sum_u64(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// gkey[i] = dim.g[row(fk[i])] and hit[i] = 1 when fk[i] has a match, else hit[i] = 0.
__global__ void __launch_bounds__(256) hk_probe_kernel(const void *fk, int fk_dtype, int64_t n_fact, long long pk_min,
                                                        long long pk_span, const uint32_t *lut, const void *g, int g_width,
                                                        void *gkey_out, uint32_t *hit_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_fact; i += stride) {
        const long long v = load_int(fk, fk_dtype, i) - pk_min;
        uint32_t row1 = 0;
        if (v >= 0 && v < pk_span) row1 = lut[v];
        hit_out[i] = row1 != 0;
        if (g_width == 4) reinterpret_cast<uint32_t *>(gkey_out)[i] = row1 ? reinterpret_cast<const uint32_t *>(g)[row1 - 1] : 0u;
        else reinterpret_cast<uint64_t *>(gkey_out)[i] = row1 ? reinterpret_cast<const uint64_t *>(g)[row1 - 1] : 0ull;
    }
}

__global__ void __launch_bounds__(256) hk_minmax_int_kernel(const void *col, int dtype, int64_t n, long long *out) {
    long long lo = LLONG_MAX, hi = LLONG_MIN;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long v = load_int(col, dtype, i);
        lo = min(lo, v);
        hi = max(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(HK_FULL_MASK, lo, o));
        hi = max(hi, __shfl_xor_sync(HK_FULL_MASK, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

unsigned grid_for(hark_ctx *ctx, int64_t n, int per_sm = 8) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * per_sm));
}

struct Bufs {
    hark_ctx *ctx;
    std::vector<void *> v;
    explicit Bufs(hark_ctx *c) : ctx(c) {}
    ~Bufs() {
        for (void *p : v) ctx->dfree(p);
    }
    int alloc(void **p, size_t bytes) {
        int rc = ctx->dalloc(p, bytes);
        if (rc == HARK_OK) v.push_back(*p);
        return rc;
    }
    void adopt(void *p) {
        if (p) v.push_back(p);
    }
};

// (key, row id) of one side, sorted by key as u32, stable
int sort_side(hark_ctx *ctx, Bufs &bufs, const hark_table *db, int32_t col, void **keys_out, void **rid_out) {
    const int64_t n = db->n;
    void *iota = nullptr;
    HK_TRY(bufs.alloc(&iota, sizeof(uint32_t) * (size_t)std::max<int64_t>(n, 1)));
    HK_TRY(hk_iota(ctx, iota, n, 4));
    std::vector<hk_sort_array> arrays(2);
    arrays[0].in = db->cols[col].ptr;
    arrays[0].width = 4;
    arrays[1].in = iota;
    arrays[1].width = 4;
    std::vector<hk_sort_keyspec> keys{hk_sort_keyspec{0, HARK_U32, 0}};
    HK_TRY(hk_radix_sort(ctx, n, keys, arrays, 0, nullptr, nullptr));
    bufs.adopt(arrays[0].result);
    bufs.adopt(arrays[1].result);
    *keys_out = arrays[0].result;
    *rid_out = arrays[1].result;
    return HARK_OK;
}

} // namespace

int hk_join(hark_ctx *ctx, hark_table **out, const hark_table *db1, const hark_table *db2, int32_t col1, int32_t col2,
            const int32_t *cols1, int64_t l, const int32_t *cols2, int64_t k) {
    const int64_t n1 = db1->n, n2 = db2->n;
    const int64_t m1 = (int64_t)db1->cols.size(), m2 = (int64_t)db2->cols.size();
    HK_ARG(ctx, l <= JMAXC && k <= JMAXC, "join: at most 16 projected columns per side");
    // join.fut:55-56 slices db1[:,col1] / db2[:,col2] whenever the table has rows
    if (n1 > 0) HK_ARG(ctx, col1 >= 0 && col1 < m1, "join: col1 out of bounds");
    if (n2 > 0) HK_ARG(ctx, col2 >= 0 && col2 < m2, "join: col2 out of bounds");
    for (auto *t : {db1, db2})
        for (auto &c : t->cols)
            HK_ARG(ctx, c.dtype == HARK_I32 || c.dtype == HARK_U32, "join: the reference entry takes u32 tables");
    ctx->entry_begin();
    std::vector<int32_t> odt((size_t)(l + k), HARK_U32);
    if (n1 == 0 || n2 == 0) {
        HK_TRY(hk_table_alloc(ctx, out, 0, 0, odt.data(), l + k));
        ctx->entry_end(0, n1 + n2, 0);
        return HARK_OK;
    }
    HK_ARG(ctx, n1 < 0xffffffffll && n2 < 0xffffffffll, "join: more than 2^32-1 rows per side is not supported");
    Bufs bufs(ctx);
    void *k1 = nullptr, *r1 = nullptr, *k2 = nullptr, *r2 = nullptr;
    HK_TRY(sort_side(ctx, bufs, db1, col1, &k1, &r1));
    HK_TRY(sort_side(ctx, bufs, db2, col2, &k2, &r2));
    uint32_t *lb = nullptr, *cnt = nullptr;
    unsigned long long *offs = nullptr;
    HK_TRY(bufs.alloc((void **)&lb, sizeof(uint32_t) * (size_t)n1));
    HK_TRY(bufs.alloc((void **)&cnt, sizeof(uint32_t) * (size_t)n1));
    HK_TRY(bufs.alloc((void **)&offs, sizeof(unsigned long long) * (size_t)(n1 + 1)));
    ctx->kernel_begin();
    hk_join_bounds_kernel<<<grid_for(ctx, n1, 16), 256, 0, ctx->stream>>>((const uint32_t *)k1, n1, (const uint32_t *)k2, n2, lb, cnt);
    HK_CHECK_LAUNCH(ctx);
    hk_join_scan_kernel<<<1, 1024, 0, ctx->stream>>>(cnt, offs, n1);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch(2);
    HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, offs + n1, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t P = (int64_t)ctx->h_scalars[0];
    if (P > 0) { // join.fut:69-73 index rows of both tables with the projected columns
        for (int64_t j = 0; j < l; j++) HK_ARG(ctx, cols1[j] >= 0 && cols1[j] < m1, "join: cols1 index out of bounds");
        for (int64_t j = 0; j < k; j++) HK_ARG(ctx, cols2[j] >= 0 && cols2[j] < m2, "join: cols2 index out of bounds");
    }
    hark_table *t = nullptr;
    HK_TRY(hk_table_alloc(ctx, &t, P, P, odt.data(), l + k));
    if (P > 0) {
        ExpandParams E;
        memset(&E, 0, sizeof E);
        E.P = P;
        E.n1 = n1;
        E.offs = offs;
        E.lb = lb;
        E.rid1 = (const uint32_t *)r1;
        E.rid2 = (const uint32_t *)r2;
        E.l = (int)l;
        E.k = (int)k;
        for (int64_t j = 0; j < l; j++) E.src1[j] = (const uint32_t *)db1->cols[cols1[j]].ptr;
        for (int64_t j = 0; j < k; j++) E.src2[j] = (const uint32_t *)db2->cols[cols2[j]].ptr;
        for (int64_t j = 0; j < l + k; j++) E.dst[j] = (uint32_t *)t->cols[j].ptr;
        hk_join_expand_kernel<<<grid_for(ctx, P, 16), 256, 0, ctx->stream>>>(E);
        cudaError_t e = cudaGetLastError();
        ctx->count_launch();
        if (e != cudaSuccess) {
            hark_table_free(ctx, t);
            return ctx->fail(HARK_ERR_CUDA, std::string("join(expand): ") + cudaGetErrorString(e));
        }
    }
    ctx->kernel_end();
    ctx->entry_end(4 * (n1 + n2) + 4 * P * (l + k) * 2, n1 + n2, P);
    *out = t;
    return HARK_OK;
}

int hk_join_groupby(hark_ctx *ctx, hark_table **out, const hark_table *fact, const hark_table *dim, int32_t fk_col,
                    int32_t pk_col, int32_t g_col, const int32_t *s_cols, const int32_t *ops, int64_t c) {
    const int64_t nf = fact->n, nd = dim->n;
    const int64_t mf = (int64_t)fact->cols.size(), md = (int64_t)dim->cols.size();
    HK_ARG(ctx, fk_col >= 0 && fk_col < mf && pk_col >= 0 && pk_col < md && g_col >= 0 && g_col < md,
           "join_groupby: column index out of bounds");
    for (int64_t j = 0; j < c; j++) HK_ARG(ctx, s_cols[j] >= 0 && s_cols[j] < mf, "join_groupby: aggregated column index out of bounds");
    HK_ARG(ctx, hk_dtype_int(fact->cols[fk_col].dtype) && hk_dtype_int(dim->cols[pk_col].dtype) &&
                    hk_dtype_int(dim->cols[g_col].dtype),
           "join_groupby: join and group keys must be integer columns");
    HK_ARG(ctx, nd < 0xffffffffll, "join_groupby: dimension table too large");
    ctx->entry_begin();
    Bufs bufs(ctx);
    const int32_t g_dtype = dim->cols[g_col].dtype;
    const int gw = hk_dtype_size(g_dtype);

    // ---- build: direct-address lookup pk -> dim row over [pk_min, pk_max] ----
    long long *d_mm = nullptr;
    HK_TRY(bufs.alloc((void **)&d_mm, 2 * sizeof(long long)));
    long long pk_min = 0, pk_span = 0;
    uint32_t *lut = nullptr;
    if (nd > 0) {
        ((long long *)ctx->h_scalars)[0] = LLONG_MAX;
        ((long long *)ctx->h_scalars)[1] = LLONG_MIN;
        HK_CUDA(ctx, cudaMemcpyAsync(d_mm, ctx->h_scalars, 2 * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        hk_minmax_int_kernel<<<grid_for(ctx, nd), 256, 0, ctx->stream>>>(dim->cols[pk_col].ptr, dim->cols[pk_col].dtype, nd, d_mm);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
        HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, d_mm, 2 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        pk_min = ((long long *)ctx->h_scalars)[0];
        const long long pk_max = ((long long *)ctx->h_scalars)[1];
        const unsigned long long span = (unsigned long long)pk_max - (unsigned long long)pk_min + 1ull;
        if (span == 0 || span > (unsigned long long)nd * 16ull + (1ull << 22))
            return ctx->fail(HARK_ERR_UNSUPPORTED, "join_groupby: dimension keys too sparse for the direct-address build "
                                                   "(span > 16 x rows); a hashed build is not implemented yet");
        pk_span = (long long)span;
        HK_TRY(bufs.alloc((void **)&lut, sizeof(uint32_t) * (size_t)span));
        HK_CUDA(ctx, cudaMemsetAsync(lut, 0, sizeof(uint32_t) * (size_t)span, ctx->stream));
        unsigned int *dup = (unsigned int *)(d_mm); // reuse: [0] as the duplicate flag

        // ---- K2 in lookup mode: probe and aggregate in one pass over the fact columns, no materialised join ----
        if (ctx->opt("groupby.impl", 0) != 1 && nf > 0 && c <= HK_DENSE_MAX_AGGS) {
            hk_dense_req rq;
            rq.n = nf;
            rq.key = fact->cols[fk_col].ptr;
            rq.key_dtype = fact->cols[fk_col].dtype;
            rq.out_key_dtype = g_dtype;
            rq.c = (int)c;
            bool eligible = true;
            std::vector<int> val_of_col((size_t)mf, -1);
            for (int64_t j = 0; j < c && eligible; j++) {
                int code = ops[j];
                if (code < HARK_AGG_PROD || code > HARK_AGG_SUMF64) code = HARK_AGG_MIN;
                rq.agg_code[j] = code;
                if (code == HARK_AGG_COUNT) {
                    rq.agg_val[j] = -1;
                    continue;
                }
                const int col = s_cols[j];
                if (val_of_col[col] < 0) {
                    if (rq.nvals == HK_DENSE_MAX_VALS) {
                        eligible = false;
                        break;
                    }
                    val_of_col[col] = rq.nvals;
                    rq.vals[rq.nvals] = fact->cols[col].ptr;
                    rq.val_dtypes[rq.nvals] = fact->cols[col].dtype;
                    rq.nvals++;
                }
                rq.agg_val[j] = val_of_col[col];
            }
            if (eligible) {
                HK_TRY(hk_column_minmax(ctx, dim->cols[g_col], nd, g_dtype, &rq.g_lo, &rq.g_hi));
                if (rq.g_hi - rq.g_lo < (1ull << 20)) {
                    unsigned long long *nz = (unsigned long long *)d_mm;
                    HK_CUDA(ctx, cudaMemsetAsync(nz, 0, sizeof(unsigned long long), ctx->stream));
                    // a lookup much larger than L2: bring the dimension rows into lookup-slice order first, so the
                    // scattered stores below complete whole sectors while the slice is still L2-resident
                    const void *pk_src = dim->cols[pk_col].ptr, *g_src = dim->cols[g_col].ptr;
                    const int pkw = hk_dtype_size(dim->cols[pk_col].dtype);
                    if ((uint64_t)span * 4ull > (96ull << 20) && gw == 4) {
                        hk_part_spec ps;
                        ps.dtype = dim->cols[pk_col].dtype;
                        ps.base = pkw == 4 ? (uint64_t)(ps.dtype == HARK_U32 ? (uint32_t)pk_min : ((uint32_t)(int32_t)pk_min ^ 0x80000000u))
                                           : ((uint64_t)pk_min ^ 0x8000000000000000ull);
                        ps.span = (uint64_t)span;
                        ps.shift = 22; // 4 Mi entries = 16 MB of lookup per bin
                        while ((((uint64_t)span - 1) >> ps.shift) + 1 > 256) ps.shift++;
                        ps.nbins = (int)((((uint64_t)span - 1) >> ps.shift) + 1);
                        void *pko = nullptr, *go[3] = {nullptr, nullptr, nullptr};
                        unsigned long long *offs = nullptr;
                        const void *gv[1] = {g_src};
                        HK_TRY(hk_partition_pass(ctx, nd, pk_src, pkw, ps, 1, gv, &pko, go, &offs));
                        bufs.adopt(pko);
                        bufs.adopt(go[0]);
                        bufs.adopt(offs);
                        pk_src = pko;
                        g_src = go[0];
                    }
                    hk_slot_lut_build_kernel<<<grid_for(ctx, nd), 256, 0, ctx->stream>>>(
                        pk_src, dim->cols[pk_col].dtype, g_src, g_dtype, nd, pk_min, (unsigned long long)rq.g_lo, lut);
                    HK_CHECK_LAUNCH(ctx);
                    hk_count_nonzero_kernel<<<grid_for(ctx, (int64_t)span / 4 + 1), 256, 0, ctx->stream>>>(lut, (int64_t)span, nz);
                    HK_CHECK_LAUNCH(ctx);
                    ctx->count_launch(2);
                    HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, nz, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
                    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                    if ((int64_t)ctx->h_scalars[0] != nd)
                        return ctx->fail(HARK_ERR_ARG, "join_groupby: dim.pk is not unique");
                    rq.lut = lut;
                    rq.pk_min = pk_min;
                    rq.pk_span = pk_span;
                    bool handled = false;
                    hark_table *t = nullptr;
                    HK_TRY(hk_dense_groupby(ctx, &t, rq, &handled));
                    if (handled) {
                        int64_t alg = nf * hk_dtype_size(rq.key_dtype) + nf * 4 * rq.nvals;
                        alg += nd * (hk_dtype_size(dim->cols[pk_col].dtype) + gw);
                        ctx->entry_end(alg, nf + nd, t->n);
                        *out = t;
                        return HARK_OK;
                    }
                    HK_CUDA(ctx, cudaMemsetAsync(lut, 0, sizeof(uint32_t) * (size_t)span, ctx->stream)); // rebuild as row lut below
                }
            }
        }
        HK_CUDA(ctx, cudaMemsetAsync(dup, 0, sizeof(unsigned int), ctx->stream));
        hk_lut_build_kernel<<<grid_for(ctx, nd), 256, 0, ctx->stream>>>(dim->cols[pk_col].ptr, dim->cols[pk_col].dtype, nd, pk_min, lut, dup);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
        HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, dup, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (((unsigned int *)ctx->h_scalars)[0] != 0)
            return ctx->fail(HARK_ERR_ARG, "join_groupby: dim.pk is not unique");
    }
    // ---- probe: group key + hit flag per fact row ----
    void *gkey = nullptr;
    uint32_t *hit = nullptr;
    HK_TRY(bufs.alloc(&gkey, (size_t)std::max<int64_t>(nf, 1) * gw));
    HK_TRY(bufs.alloc((void **)&hit, sizeof(uint32_t) * (size_t)std::max<int64_t>(nf, 1)));
    ctx->kernel_begin();
    if (nf > 0) {
        hk_probe_kernel<<<grid_for(ctx, nf, 16), 256, 0, ctx->stream>>>(fact->cols[fk_col].ptr, fact->cols[fk_col].dtype, nf, pk_min,
                                                                       pk_span, lut, dim->cols[g_col].ptr, gw, gkey, hit);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
    }
    // ---- keep the matched rows (K1), then GROUP BY the looked-up key ----
    hark_table tmp; // borrowed view: [gkey, hit, distinct fact value columns]
    tmp.n = nf;
    tmp.cap = nf;
    auto push = [&](void *p, int32_t dt) {
        hark_col col;
        col.ptr = p;
        col.dtype = dt;
        col.owned = false;
        tmp.cols.push_back(col);
    };
    push(gkey, g_dtype);
    push(hit, HARK_U32);
    std::vector<int> tmp_of_col((size_t)mf, -1);
    std::vector<int32_t> sel{0}, s2, ops2;
    for (int64_t j = 0; j < c; j++) {
        const int col = s_cols[j];
        if (tmp_of_col[col] < 0) {
            tmp_of_col[col] = (int)tmp.cols.size();
            push(fact->cols[col].ptr, fact->cols[col].dtype);
            sel.push_back(tmp_of_col[col]);
        }
    }
    for (int64_t j = 0; j < c; j++) {
        s2.push_back((int32_t)(std::find(sel.begin(), sel.end(), tmp_of_col[s_cols[j]]) - sel.begin()));
        ops2.push_back(ops[j]);
    }
    hark_pred only_hits{1, HARK_EQ, 1, 1.0};
    hark_table *matched = nullptr;
    HK_TRY(hk_filter(ctx, &matched, &tmp, sel.data(), (int64_t)sel.size(), &only_hits, 1));
    hark_table *res = nullptr;
    int rc = hk_groupby(ctx, &res, matched, 0, s2.data(), ops2.data(), c, nullptr, 0, false);
    hark_table_free(ctx, matched);
    if (rc != HARK_OK) return rc;
    int64_t alg = 0;
    alg += nf * hk_dtype_size(fact->cols[fk_col].dtype);
    for (size_t j = 2; j < tmp.cols.size(); j++) alg += nf * hk_dtype_size(tmp.cols[j].dtype);
    alg += nd * (hk_dtype_size(dim->cols[pk_col].dtype) + gw);
    ctx->entry_end(alg, nf + nd, res->n);
    *out = res;
    return HARK_OK;
}
