// filter.cu — K1: SELECT cols [WHERE p1 AND p2 ...] as one single-pass scan-filter-project-compact
// kernel, plus the predicate-free projection (query_sel) as a vectorised column copy.
//
// Reference: futhark/select.fut:9-23 (projection `map (sel cols) db`; the WHERE at :18 is a comment)
// and main.fut:7.  The reference gathers columns row by row out of a row-major [n][m] array; here
// the table is SoA, so a projection is k coalesced column copies and a filter reads only the
// predicate and selected columns (P ∪ S), never the rest of the row.
//
// Kernel shape (HBM-bound; roofline numerator = 4·N·|P∪S| + 4·N_out·|S| for 4-byte columns):
//   * persistent CTAs claim 4096-row tiles in order from a ticket counter;
//   * every thread owns 4 groups of 4 consecutive rows -> one 128-bit streaming load (ld.global.cs.v4)
//     per group and column; with static column counts all loads of a tile (predicates AND selected
//     columns) are issued before the first compare, so they are in flight across the scan;
//   * predicate -> 16-bit row mask per thread; ranks from popc + warp shuffle scan; one decoupled
//     look-back per tile gives the global output offset without a second pass over the data;
//   * selected values are staged in shared memory in output order and written back with fully
//     coalesced streaming stores, so output row order == input row order (bit-exact vs the oracle).
#include <algorithm>
#include <new>
#include <stdexcept>
#include <vector>

#include "hark_internal.cuh"

namespace {

constexpr int FT = 256;            // threads per CTA
constexpr int FG = 4;              // 4-row groups per thread
constexpr int FROWS = FG * 4;      // rows per thread
constexpr int FTILE = FT * FROWS;  // rows per tile (4096)
constexpr int FWROWS = 32 * FROWS; // rows per warp (512)
constexpr int MAXP = 8;
constexpr int MAXS = 16;

struct FilterParams {
    int np, ns;
    int64_t n;
    int64_t num_tiles;
    const void *pcol[MAXP];
    int pdtype[MAXP];
    int pop[MAXP];
    int64_t pival[MAXP];
    double pfval[MAXP];
    const void *scol[MAXS];
    void *dcol[MAXS];
    int swidth[MAXS];
    uint64_t *state;            // [num_tiles] look-back words (zeroed)
    unsigned long long *ticket; // zeroed
    unsigned long long *total;  // receives N_out
};

template <int W> struct Raw;
template <> struct Raw<4> { using T = uint32_t; };
template <> struct Raw<8> { using T = uint64_t; };

// The 16 rows of this thread for one column.  row0 = first row of group 0; group g starts 128 rows on.
template <int W>
__device__ __forceinline__ void load16(const void *col, int64_t row0, int64_t n, bool full,
                                       typename Raw<W>::T (&x)[FROWS]) {
    using T = typename Raw<W>::T;
    const T *p = reinterpret_cast<const T *>(col);
    if (full) {
#pragma unroll
        for (int g = 0; g < FG; g++) {
            if constexpr (W == 4) {
                const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(p + row0 + g * 128));
                x[g * 4 + 0] = v.x; x[g * 4 + 1] = v.y; x[g * 4 + 2] = v.z; x[g * 4 + 3] = v.w;
            } else {
                const ulonglong2 a = __ldcs(reinterpret_cast<const ulonglong2 *>(p + row0 + g * 128));
                const ulonglong2 b = __ldcs(reinterpret_cast<const ulonglong2 *>(p + row0 + g * 128 + 2));
                x[g * 4 + 0] = a.x; x[g * 4 + 1] = a.y; x[g * 4 + 2] = b.x; x[g * 4 + 3] = b.y;
            }
        }
    } else {
#pragma unroll
        for (int g = 0; g < FG; g++)
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int64_t r = row0 + g * 128 + e;
                x[g * 4 + e] = r < n ? p[r] : (T)0;
            }
    }
}

template <typename V, typename C>
__device__ __forceinline__ uint32_t cmp16(const V (&v)[FROWS], int op, C c) {
    uint32_t m = 0;
    switch (op) {
    case HARK_GT:
#pragma unroll
        for (int i = 0; i < FROWS; i++) m |= (uint32_t)(v[i] > c) << i;
        break;
    case HARK_GE:
#pragma unroll
        for (int i = 0; i < FROWS; i++) m |= (uint32_t)(v[i] >= c) << i;
        break;
    case HARK_LT:
#pragma unroll
        for (int i = 0; i < FROWS; i++) m |= (uint32_t)(v[i] < c) << i;
        break;
    case HARK_LE:
#pragma unroll
        for (int i = 0; i < FROWS; i++) m |= (uint32_t)(v[i] <= c) << i;
        break;
    case HARK_EQ:
#pragma unroll
        for (int i = 0; i < FROWS; i++) m |= (uint32_t)(v[i] == c) << i;
        break;
    default:
#pragma unroll
        for (int i = 0; i < FROWS; i++) m |= (uint32_t)(v[i] != c) << i;
        break;
    }
    return m;
}

// include/hark.h hark_pred: ints widen to int64 and compare with ival; f32 compares in f32 against
// (float)fval; f64 against fval.
template <int W>
__device__ __forceinline__ uint32_t eval_pred(const typename Raw<W>::T (&x)[FROWS], int dtype, int op, int64_t ic,
                                              double fc) {
    if constexpr (W == 4) {
        if (dtype == HARK_F32) {
            float v[FROWS];
#pragma unroll
            for (int i = 0; i < FROWS; i++) v[i] = __uint_as_float(x[i]);
            return cmp16(v, op, (float)fc);
        }
        int64_t v[FROWS];
        if (dtype == HARK_I32) {
#pragma unroll
            for (int i = 0; i < FROWS; i++) v[i] = (int64_t)(int32_t)x[i];
        } else {
#pragma unroll
            for (int i = 0; i < FROWS; i++) v[i] = (int64_t)x[i];
        }
        return cmp16(v, op, ic);
    } else {
        if (dtype == HARK_F64) {
            double v[FROWS];
#pragma unroll
            for (int i = 0; i < FROWS; i++) v[i] = __longlong_as_double((long long)x[i]);
            return cmp16(v, op, fc);
        }
        int64_t v[FROWS];
#pragma unroll
        for (int i = 0; i < FROWS; i++) v[i] = (int64_t)x[i];
        return cmp16(v, op, ic);
    }
}

// Write this thread's selected elements into the staging buffer at their tile-local output rank.
template <int W>
__device__ __forceinline__ void stage16(typename Raw<W>::T *stage, const typename Raw<W>::T (&x)[FROWS], uint32_t mask,
                                        const uint32_t (&gbase)[FG]) {
#pragma unroll
    for (int g = 0; g < FG; g++) {
        const uint32_t nib = (mask >> (g * 4)) & 0xfu;
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (nib & (1u << e)) stage[gbase[g] + __popc(nib & ((1u << e) - 1u))] = x[g * 4 + e];
    }
}

template <int W>
__device__ __forceinline__ void flush_stage(void *dst, uint64_t out_base, const typename Raw<W>::T *stage,
                                            uint32_t count) {
    using T = typename Raw<W>::T;
    T *d = reinterpret_cast<T *>(dst) + out_base;
    for (uint32_t i = threadIdx.x; i < count; i += FT) __stcs(d + i, stage[i]);
}

// W_T: 4 or 8 = every predicate and selected column has that width; 0 = mixed widths (runtime).
// NP_T / NS_T >= 0: static predicate / selected-column counts (full unrolling, all loads up front);
// -1: runtime counts.
template <int W_T, int NP_T, int NS_T>
__global__ void __launch_bounds__(FT, 2) hk_filter_kernel(const __grid_constant__ FilterParams P) {
    extern __shared__ __align__(16) unsigned char f_smem[];
    __shared__ uint32_t s_warp_tot[FT / 32];
    __shared__ unsigned long long s_excl;
    __shared__ long long s_tile;

    const int np = NP_T >= 0 ? NP_T : P.np;
    const int ns = NS_T >= 0 ? NS_T : P.ns;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr bool kStaticSel = (W_T != 0 && NS_T > 0);

    while (true) {
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(P.ticket, 1ull);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.num_tiles) break;
        const int64_t tile_base = tile * FTILE;
        const bool full = tile_base + FTILE <= P.n;
        const int64_t row0 = tile_base + (int64_t)warp * FWROWS + lane * 4;

        uint32_t mask = 0xffffu;
        if (!full) {
            mask = 0;
#pragma unroll
            for (int g = 0; g < FG; g++)
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (row0 + g * 128 + e < P.n) mask |= 1u << (g * 4 + e);
        }

        // ---- selected columns (static case): issue their loads first, they are consumed last ----
        typename Raw<(W_T ? W_T : 4)>::T xs[kStaticSel ? NS_T : 1][FROWS];
        if constexpr (kStaticSel) {
#pragma unroll
            for (int j = 0; j < NS_T; j++) load16<W_T>(P.scol[j], row0, P.n, full, xs[j]);
        }

        // ---- predicates ----
        if constexpr (NP_T >= 0 && W_T != 0) {
            typename Raw<W_T>::T xp[NP_T > 0 ? NP_T : 1][FROWS];
#pragma unroll
            for (int p = 0; p < NP_T; p++) load16<W_T>(P.pcol[p], row0, P.n, full, xp[p]);
#pragma unroll
            for (int p = 0; p < NP_T; p++) mask &= eval_pred<W_T>(xp[p], P.pdtype[p], P.pop[p], P.pival[p], P.pfval[p]);
        } else {
            for (int p = 0; p < np; p++) {
                const int dt = P.pdtype[p];
                if ((W_T == 4) || (W_T == 0 && (dt != HARK_I64 && dt != HARK_F64))) {
                    uint32_t x[FROWS];
                    load16<4>(P.pcol[p], row0, P.n, full, x);
                    mask &= eval_pred<4>(x, dt, P.pop[p], P.pival[p], P.pfval[p]);
                } else {
                    uint64_t x[FROWS];
                    load16<8>(P.pcol[p], row0, P.n, full, x);
                    mask &= eval_pred<8>(x, dt, P.pop[p], P.pival[p], P.pfval[p]);
                }
            }
        }

        // ---- tile-local ranks in row order (warp, group, lane, element) ----
        uint32_t gbase[FG];
        uint32_t warp_run = 0;
#pragma unroll
        for (int g = 0; g < FG; g++) {
            const uint32_t c = __popc((mask >> (g * 4)) & 0xfu);
            const uint32_t inc = hk_warp_incl_scan_u32(c);
            gbase[g] = warp_run + inc - c;
            warp_run += __shfl_sync(HK_FULL_MASK, inc, 31);
        }
        if (lane == 0) s_warp_tot[warp] = warp_run;
        __syncthreads();
        if (warp == 0) {
            const uint32_t t = lane < FT / 32 ? s_warp_tot[lane] : 0u;
            const uint32_t tile_cnt = hk_warp_sum_u32(t);
            const uint64_t excl = hk_lookback_u64(P.state, tile, (uint64_t)tile_cnt);
            if (lane == 0) {
                s_excl = excl;
                if (tile == P.num_tiles - 1) *P.total = excl + tile_cnt;
            }
        }
        uint32_t warp_base = 0, tile_cnt = 0;
#pragma unroll
        for (int w = 0; w < FT / 32; w++) {
            const uint32_t t = s_warp_tot[w];
            if (w < warp) warp_base += t;
            tile_cnt += t;
        }
#pragma unroll
        for (int g = 0; g < FG; g++) gbase[g] += warp_base;
        __syncthreads();
        const uint64_t out_base = s_excl;

        // ---- compaction: stage in output order, then coalesced streaming stores ----
        if constexpr (kStaticSel) {
            typename Raw<W_T>::T *stage = reinterpret_cast<typename Raw<W_T>::T *>(f_smem);
#pragma unroll
            for (int j = 0; j < NS_T; j++) {
                stage16<W_T>(stage, xs[j], mask, gbase);
                __syncthreads();
                flush_stage<W_T>(P.dcol[j], out_base, stage, tile_cnt);
                __syncthreads();
            }
        } else {
            for (int j = 0; j < ns; j++) {
                const int w = W_T ? W_T : P.swidth[j];
                if (w == 4) {
                    uint32_t x[FROWS];
                    uint32_t *stage = reinterpret_cast<uint32_t *>(f_smem);
                    load16<4>(P.scol[j], row0, P.n, full, x);
                    stage16<4>(stage, x, mask, gbase);
                    __syncthreads();
                    flush_stage<4>(P.dcol[j], out_base, stage, tile_cnt);
                } else {
                    uint64_t x[FROWS];
                    uint64_t *stage = reinterpret_cast<uint64_t *>(f_smem);
                    load16<8>(P.scol[j], row0, P.n, full, x);
                    stage16<8>(stage, x, mask, gbase);
                    __syncthreads();
                    flush_stage<8>(P.dcol[j], out_base, stage, tile_cnt);
                }
                __syncthreads();
            }
        }
    }
}

using filter_kern_t = void (*)(const FilterParams);

template <int W>
filter_kern_t pick_static(int np, int ns) {
#define HK_FK(NP, NS) \
    if (np == NP && ns == NS) return hk_filter_kernel<W, NP, NS>;
    // only the combinations ptxas fits in 128 registers without spilling (-Xptxas -v)
    HK_FK(1, 1) HK_FK(1, 2) HK_FK(2, 1)
    if constexpr (W == 4) {
        HK_FK(2, 2) HK_FK(3, 1) HK_FK(3, 2) HK_FK(1, 3) HK_FK(1, 4) HK_FK(2, 3) HK_FK(2, 4) HK_FK(3, 3)
    }
#undef HK_FK
    return hk_filter_kernel<W, -1, -1>;
}

// ---- plain column copy (projection without predicate, slices, concatenation) ----
constexpr int CP_MAXC = 16;
struct CopyParams {
    const unsigned char *src[CP_MAXC];
    unsigned char *dst[CP_MAXC];
    int64_t bytes[CP_MAXC];
};

__global__ void __launch_bounds__(256) hk_copy_kernel(const __grid_constant__ CopyParams P) {
    const int c = blockIdx.y;
    const unsigned char *s = P.src[c];
    unsigned char *d = P.dst[c];
    const int64_t bytes = P.bytes[c];
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
        const int64_t nv = bytes >> 4;
        const uint4 *sv = reinterpret_cast<const uint4 *>(s);
        uint4 *dv = reinterpret_cast<uint4 *>(d);
        int64_t i = tid;
        for (; i + 3 * nthr < nv; i += 4 * nthr) { // 4 independent 16-byte loads in flight per thread
            const uint4 a = __ldcs(sv + i), b = __ldcs(sv + i + nthr), c2 = __ldcs(sv + i + 2 * nthr),
                        e = __ldcs(sv + i + 3 * nthr);
            __stcs(dv + i, a); __stcs(dv + i + nthr, b); __stcs(dv + i + 2 * nthr, c2); __stcs(dv + i + 3 * nthr, e);
        }
        for (; i < nv; i += nthr) __stcs(dv + i, __ldcs(sv + i));
        for (int64_t b = (nv << 4) + tid; b < bytes; b += nthr) d[b] = s[b];
    } else { // unaligned slice: 4-byte words (every column element is 4 or 8 bytes)
        const int64_t nw = bytes >> 2;
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(s);
        uint32_t *dw = reinterpret_cast<uint32_t *>(d);
        for (int64_t i = tid; i < nw; i += nthr) dw[i] = sw[i];
    }
}

} // namespace

int hk_copy_columns(hark_ctx *ctx, hark_table *dst, const hark_table *src, const int32_t *cols, int64_t k,
                    int64_t src_row0, int64_t nrows, int64_t dst_row0) {
    if (nrows == 0 || k == 0) return HARK_OK;
    for (int64_t j0 = 0; j0 < k; j0 += CP_MAXC) {
        const int kc = (int)std::min<int64_t>(CP_MAXC, k - j0);
        CopyParams P;
        int64_t maxb = 0;
        for (int j = 0; j < kc; j++) {
            const hark_col &sc = src->cols[cols[j0 + j]];
            const int w = hk_dtype_size(sc.dtype);
            P.src[j] = (const unsigned char *)sc.ptr + (size_t)src_row0 * w;
            P.dst[j] = (unsigned char *)dst->cols[j0 + j].ptr + (size_t)dst_row0 * w;
            P.bytes[j] = nrows * w;
            maxb = std::max(maxb, P.bytes[j]);
        }
        const int64_t want = (maxb / 16 + 255) / 256 / 4 + 1;
        const unsigned gx = (unsigned)std::min<int64_t>(want, (int64_t)ctx->num_sms * 8);
        hk_copy_kernel<<<dim3(gx, kc), 256, 0, ctx->stream>>>(P);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
    }
    return HARK_OK;
}

int hk_filter(hark_ctx *ctx, hark_table **out, const hark_table *db, const int32_t *cols, int64_t k,
              const hark_pred *preds, int64_t np) {
    const int64_t n = db->n, m = (int64_t)db->cols.size();
    for (int64_t j = 0; j < k; j++)
        HK_ARG(ctx, cols[j] >= 0 && cols[j] < m, "query: selected column index out of bounds");
    for (int64_t p = 0; p < np; p++) {
        HK_ARG(ctx, preds[p].col >= 0 && preds[p].col < m, "query: predicate column index out of bounds");
        HK_ARG(ctx, preds[p].op >= HARK_GT && preds[p].op <= HARK_NE, "query: bad comparison operator");
    }
    HK_ARG(ctx, np <= MAXP, "query: at most 8 conjuncts are supported");
    std::vector<int32_t> dts;
    for (int64_t j = 0; j < k; j++) dts.push_back(db->cols[cols[j]].dtype);

    ctx->entry_begin();
    hark_table *t = nullptr;
    HK_TRY(hk_table_alloc(ctx, &t, np == 0 ? n : 0, n, dts.data(), k));

    int64_t alg_bytes = 0;
    if (np == 0) { // select.fut:19 — projection only
        ctx->kernel_begin();
        int rc = hk_copy_columns(ctx, t, db, cols, k, 0, n, 0);
        ctx->kernel_end();
        if (rc != HARK_OK) {
            hark_table_free(ctx, t);
            return rc;
        }
        for (int64_t j = 0; j < k; j++) alg_bytes += 2 * n * hk_dtype_size(dts[j]);
        ctx->entry_end(alg_bytes, n, n);
        *out = t;
        return HARK_OK;
    }
    if (n == 0) {
        ctx->entry_end(0, 0, 0);
        *out = t;
        return HARK_OK;
    }

    const int64_t num_tiles = (n + FTILE - 1) / FTILE;
    uint64_t *scratch = nullptr; // [0] ticket, [1] total, [2..] tile states
    const size_t scratch_bytes = (size_t)(num_tiles + 2) * sizeof(uint64_t);
    int rc = ctx->dalloc((void **)&scratch, scratch_bytes);
    if (rc != HARK_OK) {
        hark_table_free(ctx, t);
        return rc;
    }

    // widths: one static width if every involved column agrees
    int wall = 0;
    bool mixed = false;
    auto see = [&](int dt) {
        const int w = hk_dtype_size(dt);
        if (wall == 0) wall = w;
        else if (wall != w) mixed = true;
    };
    for (int64_t p = 0; p < np; p++) see(db->cols[preds[p].col].dtype);
    for (int64_t j = 0; j < k; j++) see(dts[j]);

    cudaError_t e = cudaSuccess;
    const int64_t impl = ctx->opt("filter.impl", 0); // 0 auto, 1 force the runtime-count kernel
    uint64_t n_out = 0;
    // more than MAXS selected columns: several launches over column groups (predicates re-evaluated)
    for (int64_t j0 = 0; j0 < std::max<int64_t>(k, 1) && e == cudaSuccess; j0 += MAXS) {
        const int ks = (int)std::min<int64_t>(MAXS, k - j0);
        FilterParams P;
        memset(&P, 0, sizeof P);
        P.np = (int)np;
        P.ns = ks;
        P.n = n;
        P.num_tiles = num_tiles;
        for (int64_t p = 0; p < np; p++) {
            const hark_col &c = db->cols[preds[p].col];
            P.pcol[p] = c.ptr;
            P.pdtype[p] = c.dtype;
            P.pop[p] = preds[p].op;
            P.pival[p] = preds[p].ival;
            P.pfval[p] = preds[p].fval;
        }
        for (int j = 0; j < ks; j++) {
            P.scol[j] = db->cols[cols[j0 + j]].ptr;
            P.dcol[j] = t->cols[j0 + j].ptr;
            P.swidth[j] = hk_dtype_size(dts[j0 + j]);
        }
        P.ticket = (unsigned long long *)scratch;
        P.total = (unsigned long long *)(scratch + 1);
        P.state = scratch + 2;
        e = cudaMemsetAsync(scratch, 0, scratch_bytes, ctx->stream);
        if (e != cudaSuccess) break;

        filter_kern_t kern;
        int wsm;
        if (mixed) {
            kern = hk_filter_kernel<0, -1, -1>;
            wsm = 8;
        } else if (wall == 4) {
            kern = impl == 1 ? hk_filter_kernel<4, -1, -1> : pick_static<4>((int)np, ks);
            wsm = 4;
        } else {
            kern = impl == 1 ? hk_filter_kernel<8, -1, -1> : pick_static<8>((int)np, ks);
            wsm = 8;
        }
        const size_t smem = (size_t)FTILE * wsm;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FT, smem);
        if (e != cudaSuccess) break;
        const int64_t want_occ = ctx->opt("filter.ctas_per_sm", 0);
        if (want_occ > 0) occ = (int)std::min<int64_t>(occ, want_occ);
        occ = std::max(occ, 1);
        const unsigned grid = (unsigned)std::min<int64_t>(num_tiles, (int64_t)ctx->num_sms * occ);
        if (j0 == 0) ctx->kernel_begin();
        kern<<<grid, FT, smem, ctx->stream>>>(P);
        e = cudaGetLastError();
        ctx->count_launch();
        if (j0 + MAXS >= k) ctx->kernel_end();
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ctx->h_scalars, scratch + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ctx->dfree(scratch);
    if (e != cudaSuccess) {
        hark_table_free(ctx, t);
        return ctx->fail(HARK_ERR_CUDA, std::string("query_filter: ") + cudaGetErrorString(e));
    }
    n_out = ctx->h_scalars[0];
    t->n = (int64_t)n_out;

    // algorithmic bytes: each distinct involved column read once + each output column written once
    std::vector<int> seen((size_t)m, 0);
    for (int64_t p = 0; p < np; p++) seen[preds[p].col] = 1;
    for (int64_t j = 0; j < k; j++) seen[cols[j]] = 1;
    for (int64_t c = 0; c < m; c++)
        if (seen[c]) alg_bytes += n * hk_dtype_size(db->cols[c].dtype);
    for (int64_t j = 0; j < k; j++) alg_bytes += (int64_t)n_out * hk_dtype_size(dts[j]);
    ctx->entry_end(alg_bytes, n, (int64_t)n_out);
    *out = t;
    return HARK_OK;
}
