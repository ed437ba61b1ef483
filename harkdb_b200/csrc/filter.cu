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
#include "sort.cuh"

namespace {

constexpr int FT = 256;            // threads per CTA
constexpr int FG = 4;              // 4-row groups per thread
constexpr int FROWS = FG * 4;      // rows per thread
constexpr int FTILE = FT * FROWS;  // rows per tile (4096)
constexpr int FWROWS = 32 * FROWS; // rows per warp (512)
constexpr int MAXP = 16;
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

template <int N, typename V, typename C>
__device__ __forceinline__ uint32_t cmpN(const V (&v)[N], int op, C c) {
    uint32_t m = 0;
    switch (op) {
    case HARK_GT:
#pragma unroll
        for (int i = 0; i < N; i++) m |= (uint32_t)(v[i] > c) << i;
        break;
    case HARK_GE:
#pragma unroll
        for (int i = 0; i < N; i++) m |= (uint32_t)(v[i] >= c) << i;
        break;
    case HARK_LT:
#pragma unroll
        for (int i = 0; i < N; i++) m |= (uint32_t)(v[i] < c) << i;
        break;
    case HARK_LE:
#pragma unroll
        for (int i = 0; i < N; i++) m |= (uint32_t)(v[i] <= c) << i;
        break;
    case HARK_EQ:
#pragma unroll
        for (int i = 0; i < N; i++) m |= (uint32_t)(v[i] == c) << i;
        break;
    default:
#pragma unroll
        for (int i = 0; i < N; i++) m |= (uint32_t)(v[i] != c) << i;
        break;
    }
    return m;
}

// include/hark.h hark_pred: ints widen to int64 and compare with ival; f32 compares in f32 against
// (float)fval; f64 against fval.
template <int W, int N>
__device__ __forceinline__ uint32_t eval_predN(const typename Raw<W>::T (&x)[N], int dtype, int op, int64_t ic,
                                               double fc) {
    if constexpr (W == 4) {
        if (dtype == HARK_F32) {
            float v[N];
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = __uint_as_float(x[i]);
            return cmpN<N>(v, op, (float)fc);
        }
        int64_t v[N];
        if (dtype == HARK_I32) {
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = (int64_t)(int32_t)x[i];
        } else {
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = (int64_t)x[i];
        }
        return cmpN<N>(v, op, ic);
    } else {
        if (dtype == HARK_F64) {
            double v[N];
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = __longlong_as_double((long long)x[i]);
            return cmpN<N>(v, op, fc);
        }
        int64_t v[N];
#pragma unroll
        for (int i = 0; i < N; i++) v[i] = (int64_t)x[i];
        return cmpN<N>(v, op, ic);
    }
}

template <int W>
__device__ __forceinline__ uint32_t eval_pred(const typename Raw<W>::T (&x)[FROWS], int dtype, int op, int64_t ic,
                                              double fc) {
    return eval_predN<W, FROWS>(x, dtype, op, ic, fc);
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


// ==========================================================================================
// K1 v2 — two-phase super-tile kernel (default).
//
// ncu on v1 (profiles/r01_filter_v1_ncu.txt): DRAM traffic == algorithmic bytes, but only 40 % of the
// measured copy bandwidth — 18 of 28 stall cycles per issue were the block barriers around the per-tile
// scan/look-back/staging.  v2 removes every block barrier from the data loops:
//   * a CTA claims a SUPER-TILE of 65536 rows (ticket order); warp w owns rows [w*8192, (w+1)*8192) of it;
//   * phase A: each warp streams its rows of the PREDICATE columns (batches of 1024 rows, up to 8 x 128-bit
//     loads per lane and column in flight) and parks the result as one 32-bit row mask per lane and batch;
//   * one look-back per super-tile (15 K look-backs for 1e9 rows instead of 244 K): 2 block barriers;
//   * phase B: each warp streams its rows of the SELECTED columns, loading only 16-byte groups that contain
//     a selected row, compacts through a warp-private staging buffer (__syncwarp only) and writes
//     coalesced streaming stores at its own output offset.
// Every column is still read at most once; output order == input order.
// The first cut of v2 (profiles/r01_filter_v2a_ncu.txt) reached 76 % of the measured peak and stalled on
// `no_instruction`: 20 K SASS instructions from unrolled batches and a 6-way operator switch around every
// compare.  Hence: predicates are canonicalised on the host into a branch-free form (three enable bits +
// invert), batch loops are real loops (masks parked in shared memory), and the tail needs no per-element
// guarded loads because every table allocation is padded to 256 bytes.
// ==========================================================================================
#ifndef F2_NB_DEF
#define F2_NB_DEF 8
#endif
#ifndef F2_MINBLOCKS
#define F2_MINBLOCKS 2
#endif
#ifndef F2_UL_CAP
#define F2_UL_CAP 4   // groups per load wave (static kernels); two waves are kept in flight (software pipeline)
#endif
constexpr int F2_U = 8;                         // 4-row groups per lane and batch -> 32 rows -> one 32-bit mask
constexpr int F2_NB = F2_NB_DEF;                // batches per warp and super-tile
constexpr int F2_BROWS = 32 * 4 * F2_U;         // rows per warp batch (1024)
constexpr int F2_WROWS = F2_BROWS * F2_NB;      // rows per warp (4096)
constexpr int F2_TILE = F2_WROWS * (FT / 32);   // rows per super-tile (32768)

// Canonical predicate: result = ((g && x > c) || (e && x == c) || (l && x < c)) ^ inv.
// cls 0: IEEE compare of floats (c = bit pattern).  cls 1: unsigned compare of (x ^ bias) against c, which is
// the signed order when bias = sign bit.  NE = EQ with inv (so NaN != c is true); constants outside a 32-bit
// column's range fold to en = 7 (always) or en = 0 (never) on the host.
struct CanonPred {
    uint64_t c;
    uint64_t bias;
    int cls;
    int en; // bit0 gt, bit1 eq, bit2 lt
    int inv;
    int width;
    int clause_end; // 1: last predicate of its OR-clause (a plain conjunct is a clause of one)
};

struct Filter2Params {
    int np, ns;
    int64_t n;
    int64_t num_tiles;
    const void *pcol[MAXP];
    CanonPred pred[MAXP];
    const void *scol[MAXS];
    void *dcol[MAXS];
    int swidth[MAXS];
    uint64_t *state;
    unsigned long long *ticket;
    unsigned long long *total;
};

// groups per load wave so that one wave of `cnt` columns keeps <= 64 data registers per lane
constexpr int f2_ul(int cnt, int wf) {
    const int v0 = 16 / ((cnt > 0 ? cnt : 1) * wf);
    const int v = v0 > F2_UL_CAP ? F2_UL_CAP : v0;
    return v >= 8 ? 8 : v >= 4 ? 4 : v >= 2 ? 2 : 1;
}
constexpr int f2_min(int a, int b) { return a < b ? a : b; }

// Loads groups [u0, u0+UL) of this lane's batch rows (lrow = batch_row0 + lane*4; group u is 128 rows on).
// A group is loaded when its nibble in `need` is non-zero and its first row is < n.  The 16/32-byte load of a
// partially valid last group stays inside the allocation: bases are 16-byte aligned and every allocation is
// padded to 256 bytes (hark_ctx::dalloc); rows >= n are masked out by the caller.
template <int W, int UL>
__device__ __forceinline__ void f2_load_wave(const void *col, int64_t lrow, int u0, int64_t n, uint32_t need,
                                             typename Raw<W>::T (&x)[UL * 4]) {
    using T = typename Raw<W>::T;
    const T *p = reinterpret_cast<const T *>(col);
#pragma unroll
    for (int k = 0; k < UL; k++) {
        const int64_t r = lrow + (int64_t)(u0 + k) * 128;
        const bool want = (((need >> ((u0 + k) * 4)) & 0xfu) != 0) && r < n;
        if constexpr (W == 4) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (want) v = __ldcs(reinterpret_cast<const uint4 *>(p + r));
            x[k * 4 + 0] = v.x; x[k * 4 + 1] = v.y; x[k * 4 + 2] = v.z; x[k * 4 + 3] = v.w;
        } else {
            ulonglong2 a = make_ulonglong2(0, 0), b = make_ulonglong2(0, 0);
            if (want) {
                a = __ldcs(reinterpret_cast<const ulonglong2 *>(p + r));
                b = __ldcs(reinterpret_cast<const ulonglong2 *>(p + r + 2));
            }
            x[k * 4 + 0] = a.x; x[k * 4 + 1] = a.y; x[k * 4 + 2] = b.x; x[k * 4 + 3] = b.y;
        }
    }
}

template <int W, int N>
__device__ __forceinline__ uint32_t f2_eval(const typename Raw<W>::T (&x)[N], const CanonPred &q) {
    const bool g = q.en & 1, e = q.en & 2, l = q.en & 4;
    uint32_t m = 0;
    if (q.cls == 0) {
        if constexpr (W == 4) {
            const float c = __uint_as_float((uint32_t)q.c);
#pragma unroll
            for (int i = 0; i < N; i++) {
                const float v = __uint_as_float(x[i]);
                m |= (uint32_t)((g && v > c) || (e && v == c) || (l && v < c)) << i;
            }
        } else {
            const double c = __longlong_as_double((long long)q.c);
#pragma unroll
            for (int i = 0; i < N; i++) {
                const double v = __longlong_as_double((long long)x[i]);
                m |= (uint32_t)((g && v > c) || (e && v == c) || (l && v < c)) << i;
            }
        }
    } else {
        using T = typename Raw<W>::T;
        const T c = (T)q.c, bias = (T)q.bias;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const T v = x[i] ^ bias;
            m |= (uint32_t)((g && v > c) || (e && v == c) || (l && v < c)) << i;
        }
    }
    constexpr uint32_t all = (N == 32) ? 0xffffffffu : ((1u << N) - 1u);
    return q.inv ? (~m & all) : m;
}

// AND the wave's predicate result into the batch mask (bits [u0*4, (u0+UL)*4)).
template <int UL>
__device__ __forceinline__ uint32_t f2_merge(uint32_t mask, uint32_t sub, int u0) {
    constexpr uint32_t wave_bits = (UL * 4 == 32) ? 0xffffffffu : ((1u << (UL * 4)) - 1u);
    return mask & ((sub << (u0 * 4)) | ~(wave_bits << (u0 * 4)));
}

// Stage the selected elements of one wave of one column (warp-private buffer, slot = rank inside the wave).
template <int W, int UL>
__device__ __forceinline__ void f2_stage_wave(typename Raw<W>::T *stage, const typename Raw<W>::T (&x)[UL * 4],
                                              uint32_t mask, int u0, const uint32_t (&gb)[F2_U], uint32_t wave_first) {
#pragma unroll
    for (int k = 0; k < UL; k++) {
        const uint32_t nib = (mask >> ((u0 + k) * 4)) & 0xfu;
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (nib & (1u << e)) stage[gb[u0 + k] - wave_first + __popc(nib & ((1u << e) - 1u))] = x[k * 4 + e];
    }
}

template <int W>
__device__ __forceinline__ void f2_flush(void *dst, uint64_t off, const typename Raw<W>::T *stage, uint32_t cnt) {
    using T = typename Raw<W>::T;
    T *d = reinterpret_cast<T *>(dst) + off;
    for (uint32_t i = threadIdx.x & 31; i < cnt; i += 32) __stcs(d + i, stage[i]);
}

template <int W_T, int NP_T, int NS_T>
__global__ void __launch_bounds__(FT, F2_MINBLOCKS) hk_filter2_kernel(const __grid_constant__ Filter2Params P) {
    extern __shared__ __align__(16) unsigned char f_smem[];
    __shared__ uint32_t s_mask[F2_NB][FT];
    __shared__ uint32_t s_warp_tot[FT / 32];
    __shared__ unsigned long long s_excl;
    __shared__ long long s_tile;

    constexpr bool kStatic = (W_T != 0 && NP_T > 0 && NS_T > 0);
    constexpr int WF = (W_T == 8) ? 2 : 1;
    constexpr int UL = kStatic ? f2_min(f2_ul(NP_T, WF), f2_ul(NS_T, WF)) : 4;
    constexpr int NSS = kStatic ? NS_T : 1;
    // warp-private staging: NSS regions of UL*128 slots (static) / one region of 512 8-byte slots (dynamic)
    constexpr int STAGE_BYTES_PER_WARP = kStatic ? NSS * UL * 128 * W_T : 4 * 128 * 8;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wstage = f_smem + (size_t)warp * STAGE_BYTES_PER_WARP;
    const int np = P.np, ns = P.ns;

    while (true) {
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(P.ticket, 1ull);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.num_tiles) break;
        const int64_t wrow0 = tile * F2_TILE + (int64_t)warp * F2_WROWS;

        // ---------------- phase A: predicate columns -> row masks ----------------
        uint32_t wtot = 0;
#pragma unroll 1
        for (int b = 0; b < F2_NB; b++) {
            const int64_t brow0 = wrow0 + (int64_t)b * F2_BROWS;
            const int64_t lrow = brow0 + lane * 4;
            uint32_t mask = 0;
            if (brow0 < P.n) { // warp-uniform
                mask = 0xffffffffu;
                if (brow0 + F2_BROWS > P.n) { // ragged end: clear rows >= n
                    mask = 0;
#pragma unroll
                    for (int u = 0; u < F2_U; u++) {
                        const int64_t left = P.n - (lrow + u * 128);
                        const uint32_t nib = left >= 4 ? 0xfu : left <= 0 ? 0u : ((1u << (int)left) - 1u);
                        mask |= nib << (u * 4);
                    }
                }
                if constexpr (kStatic) {
                    // software pipeline: the loads of wave k+1 are issued before wave k is evaluated
                    typename Raw<W_T>::T xa[NP_T][UL * 4], xb[NP_T][UL * 4];
#pragma unroll
                    for (int p = 0; p < NP_T; p++) f2_load_wave<W_T, UL>(P.pcol[p], lrow, 0, P.n, 0xffffffffu, xa[p]);
#pragma unroll
                    for (int u0 = 0; u0 < F2_U; u0 += 2 * UL) {
                        if (u0 + UL < F2_U) {
#pragma unroll
                            for (int p = 0; p < NP_T; p++) f2_load_wave<W_T, UL>(P.pcol[p], lrow, u0 + UL, P.n, 0xffffffffu, xb[p]);
                        }
#pragma unroll
                        for (int p = 0; p < NP_T; p++) mask = f2_merge<UL>(mask, f2_eval<W_T, UL * 4>(xa[p], P.pred[p]), u0);
                        if (u0 + 2 * UL < F2_U) {
#pragma unroll
                            for (int p = 0; p < NP_T; p++) f2_load_wave<W_T, UL>(P.pcol[p], lrow, u0 + 2 * UL, P.n, 0xffffffffu, xa[p]);
                        }
                        if (u0 + UL < F2_U) {
#pragma unroll
                            for (int p = 0; p < NP_T; p++) mask = f2_merge<UL>(mask, f2_eval<W_T, UL * 4>(xb[p], P.pred[p]), u0 + UL);
                        }
                    }
                } else {
                    // WHERE in conjunctive normal form: the batch mask is the AND over clauses of the OR over each
                    // clause's predicates (hark.h HARK_PRED_OR); `cl` collects the clause being evaluated
                    uint32_t cl = 0;
#pragma unroll 1
                    for (int p = 0; p < np; p++) {
                        const bool w4 = (W_T == 4) || (W_T == 0 && P.pred[p].width == 4);
#pragma unroll
                        for (int u0 = 0; u0 < F2_U; u0 += UL) {
                            if (w4) {
                                uint32_t x[UL * 4];
                                f2_load_wave<4, UL>(P.pcol[p], lrow, u0, P.n, 0xffffffffu, x);
                                cl |= f2_eval<4, UL * 4>(x, P.pred[p]) << (u0 * 4);
                            } else {
                                uint64_t x[UL * 4];
                                f2_load_wave<8, UL>(P.pcol[p], lrow, u0, P.n, 0xffffffffu, x);
                                cl |= f2_eval<8, UL * 4>(x, P.pred[p]) << (u0 * 4);
                            }
                        }
                        if (P.pred[p].clause_end) {
                            mask &= cl;
                            cl = 0;
                        }
                    }
                }
            }
            s_mask[b][threadIdx.x] = mask; // read back only by this thread
            wtot += __popc(mask);
        }
        wtot = hk_warp_sum_u32(wtot);

        // ---------------- one chained-scan step per super-tile ----------------
        if (lane == 0) s_warp_tot[warp] = wtot;
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
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < FT / 32; w++)
            if (w < warp) wbase += s_warp_tot[w];
        __syncthreads();
        uint64_t off = s_excl + wbase;

        // ---------------- phase B: selected columns -> compacted output ----------------
        if (ns > 0 && wtot > 0) {
#pragma unroll 1
            for (int b = 0; b < F2_NB; b++) {
                const uint32_t mask = s_mask[b][threadIdx.x];
                if (__ballot_sync(HK_FULL_MASK, mask != 0) == 0) continue; // warp-uniform
                const int64_t lrow = wrow0 + (int64_t)b * F2_BROWS + lane * 4;
                uint32_t gb[F2_U], runb[F2_U + 1];
                uint32_t run = 0;
#pragma unroll
                for (int u = 0; u < F2_U; u++) {
                    const uint32_t c = __popc((mask >> (u * 4)) & 0xfu);
                    const uint32_t inc = hk_warp_incl_scan_u32(c);
                    runb[u] = run;
                    gb[u] = run + inc - c;
                    run += __shfl_sync(HK_FULL_MASK, inc, 31);
                }
                runb[F2_U] = run;
                if constexpr (kStatic) {
                    using T = typename Raw<W_T>::T;
                    T xa[NS_T][UL * 4], xb[NS_T][UL * 4];
                    auto consume = [&](T (&x)[NS_T][UL * 4], int u0) {
                        const uint32_t wave_first = runb[u0], wave_cnt = runb[u0 + UL] - runb[u0];
                        if (wave_cnt == 0) return; // warp-uniform
#pragma unroll
                        for (int j = 0; j < NS_T; j++)
                            f2_stage_wave<W_T, UL>(reinterpret_cast<T *>(wstage) + j * (UL * 128), x[j], mask, u0, gb, wave_first);
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < NS_T; j++)
                            f2_flush<W_T>(P.dcol[j], off + wave_first, reinterpret_cast<T *>(wstage) + j * (UL * 128), wave_cnt);
                        __syncwarp();
                    };
#pragma unroll
                    for (int j = 0; j < NS_T; j++) f2_load_wave<W_T, UL>(P.scol[j], lrow, 0, P.n, mask, xa[j]);
#pragma unroll
                    for (int u0 = 0; u0 < F2_U; u0 += 2 * UL) {
                        if (u0 + UL < F2_U) {
#pragma unroll
                            for (int j = 0; j < NS_T; j++) f2_load_wave<W_T, UL>(P.scol[j], lrow, u0 + UL, P.n, mask, xb[j]);
                        }
                        consume(xa, u0);
                        if (u0 + 2 * UL < F2_U) {
#pragma unroll
                            for (int j = 0; j < NS_T; j++) f2_load_wave<W_T, UL>(P.scol[j], lrow, u0 + 2 * UL, P.n, mask, xa[j]);
                        }
                        if (u0 + UL < F2_U) consume(xb, u0 + UL);
                    }
                } else {
#pragma unroll
                    for (int u0 = 0; u0 < F2_U; u0 += UL) {
                        const uint32_t wave_first = runb[u0], wave_cnt = runb[u0 + UL] - runb[u0];
                        if (wave_cnt == 0) continue; // warp-uniform
#pragma unroll 1
                        for (int j = 0; j < ns; j++) {
                            const int w = W_T ? W_T : P.swidth[j];
                            if (w == 4) {
                                uint32_t x[UL * 4];
                                f2_load_wave<4, UL>(P.scol[j], lrow, u0, P.n, mask, x);
                                f2_stage_wave<4, UL>(reinterpret_cast<uint32_t *>(wstage), x, mask, u0, gb, wave_first);
                                __syncwarp();
                                f2_flush<4>(P.dcol[j], off + wave_first, reinterpret_cast<uint32_t *>(wstage), wave_cnt);
                            } else {
                                uint64_t x[UL * 4];
                                f2_load_wave<8, UL>(P.scol[j], lrow, u0, P.n, mask, x);
                                f2_stage_wave<8, UL>(reinterpret_cast<uint64_t *>(wstage), x, mask, u0, gb, wave_first);
                                __syncwarp();
                                f2_flush<8>(P.dcol[j], off + wave_first, reinterpret_cast<uint64_t *>(wstage), wave_cnt);
                            }
                            __syncwarp();
                        }
                    }
                }
                off += run;
            }
        }
    }
}

template <int W_T, int NP_T, int NS_T>
constexpr size_t f2_smem_bytes() {
    constexpr bool kStatic = (W_T != 0 && NP_T > 0 && NS_T > 0);
    constexpr int WF = (W_T == 8) ? 2 : 1;
    constexpr int UL = kStatic ? f2_min(f2_ul(NP_T, WF), f2_ul(NS_T, WF)) : 4;
    return (size_t)(FT / 32) * (kStatic ? (size_t)NS_T * UL * 128 * W_T : (size_t)4 * 128 * 8);
}

struct f2_choice {
    void (*kern)(const Filter2Params);
    size_t smem;
};

template <int W>
f2_choice pick_static2(int np, int ns) {
#define HK_F2(NP, NS) \
    if (np == NP && ns == NS) return {hk_filter2_kernel<W, NP, NS>, f2_smem_bytes<W, NP, NS>()};
    HK_F2(1, 1) HK_F2(1, 2) HK_F2(2, 1) HK_F2(2, 2)
    if constexpr (W == 4) {
        HK_F2(3, 1) HK_F2(3, 2) HK_F2(1, 3) HK_F2(1, 4) HK_F2(2, 3) HK_F2(2, 4) HK_F2(3, 3) HK_F2(3, 4)
    }
#undef HK_F2
    return {hk_filter2_kernel<W, -1, -1>, f2_smem_bytes<W, -1, -1>()};
}

// include/hark.h hark_pred -> CanonPred (see the struct comment)
CanonPred canonicalise(const hark_pred &pr, int dtype) {
    CanonPred q;
    memset(&q, 0, sizeof q);
    q.width = hk_dtype_size(dtype);
    static const int en_of[6] = {1, 3, 4, 6, 2, 2}; // GT GE LT LE EQ NE(=EQ ^ inv)
    const int op = pr.op & HARK_PRED_OP_MASK;
    const int neg = (pr.op & HARK_PRED_NOT) ? 1 : 0;
    q.clause_end = (pr.op & HARK_PRED_OR) ? 0 : 1;
    q.en = en_of[op];
    q.inv = (op == HARK_NE) ^ neg;
    if (dtype == HARK_F32) {
        const float f = (float)pr.fval;
        uint32_t bits;
        memcpy(&bits, &f, 4);
        q.c = bits;
        q.cls = 0;
        return q;
    }
    if (dtype == HARK_F64) {
        memcpy(&q.c, &pr.fval, 8);
        q.cls = 0;
        return q;
    }
    q.cls = 1;
    int64_t lo, hi;
    if (dtype == HARK_I32) lo = INT32_MIN, hi = INT32_MAX, q.bias = 0x80000000ull;
    else if (dtype == HARK_U32) lo = 0, hi = UINT32_MAX, q.bias = 0;
    else lo = INT64_MIN, hi = INT64_MAX, q.bias = 0x8000000000000000ull;
    const int64_t c = pr.ival;
    if (c < lo || c > hi) { // constant outside the column's range: the comparison is a constant
        const bool below = c < lo; // every x is > c
        bool always;
        switch (op) {
        case HARK_GT: case HARK_GE: always = below; break;
        case HARK_LT: case HARK_LE: always = !below; break;
        case HARK_EQ: always = false; break;
        default: always = true; break;
        }
        q.en = (always ^ (neg != 0)) ? 7 : 0;
        q.inv = 0;
        q.c = 0;
        return q;
    }
    if (q.width == 4) q.c = ((uint64_t)(uint32_t)c) ^ q.bias;
    else q.c = (uint64_t)c ^ q.bias;
    return q;
}

using filter_kern_t = void (*)(const FilterParams);

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

int hk_copy_bytes(hark_ctx *ctx, void *dst, const void *src, int64_t bytes) {
    if (bytes <= 0) return HARK_OK;
    CopyParams P;
    P.src[0] = (const unsigned char *)src;
    P.dst[0] = (unsigned char *)dst;
    P.bytes[0] = bytes;
    const int64_t want = (bytes / 16 + 255) / 256 / 4 + 1;
    const unsigned gx = (unsigned)std::min<int64_t>(want, (int64_t)ctx->num_sms * 8);
    hk_copy_kernel<<<dim3(gx, 1), 256, 0, ctx->stream>>>(P);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch();
    return HARK_OK;
}

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
        HK_ARG(ctx, (preds[p].op & ~(HARK_PRED_OP_MASK | HARK_PRED_OR | HARK_PRED_NOT)) == 0 &&
                        (preds[p].op & HARK_PRED_OP_MASK) <= HARK_NE, "query: bad comparison operator");
    }
    HK_ARG(ctx, np <= MAXP, "query: at most 16 predicates are supported");
    HK_ARG(ctx, np == 0 || !(preds[np - 1].op & HARK_PRED_OR), "query: the last predicate cannot be OR-ed with a next one");
    bool cnf = false; // any OR / NOT: the runtime-count v2 kernel evaluates clauses; v1 and the static kernels are AND-only
    for (int64_t p = 0; p < np; p++) cnf = cnf || (preds[p].op & (HARK_PRED_OR | HARK_PRED_NOT));
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
            hk_table_free(ctx, t);
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
        hk_table_free(ctx, t);
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
    const int64_t impl = cnf ? 3 : ctx->opt("filter.impl", 0);
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

        // filter.impl: 0 = v2 two-phase super-tile kernel, static counts (default); 3 = v2 with runtime counts;
        // 1 = v1 per-tile kernel (kept as an independent second implementation for A/B parity tests)
        size_t smem;
        int64_t tiles_this = num_tiles;
        const void *kern_ptr;
        filter_kern_t kern1 = nullptr;
        f2_choice ch{nullptr, 0};
        Filter2Params P2;
        if (impl == 1) {
            kern1 = mixed ? hk_filter_kernel<0, -1, -1> : wall == 4 ? hk_filter_kernel<4, -1, -1> : hk_filter_kernel<8, -1, -1>;
            smem = (size_t)FTILE * (mixed || wall == 8 ? 8 : 4);
            kern_ptr = (const void *)kern1;
        } else {
            if (mixed) ch = {hk_filter2_kernel<0, -1, -1>, f2_smem_bytes<0, -1, -1>()};
            else if (wall == 4) ch = impl == 3 ? f2_choice{hk_filter2_kernel<4, -1, -1>, f2_smem_bytes<4, -1, -1>()} : pick_static2<4>((int)np, ks);
            else ch = impl == 3 ? f2_choice{hk_filter2_kernel<8, -1, -1>, f2_smem_bytes<8, -1, -1>()} : pick_static2<8>((int)np, ks);
            smem = ch.smem;
            kern_ptr = (const void *)ch.kern;
            tiles_this = (n + F2_TILE - 1) / F2_TILE;
            memset(&P2, 0, sizeof P2);
            P2.np = P.np; P2.ns = P.ns; P2.n = n; P2.num_tiles = tiles_this;
            for (int64_t p = 0; p < np; p++) {
                P2.pcol[p] = P.pcol[p];
                P2.pred[p] = canonicalise(preds[p], db->cols[preds[p].col].dtype);
            }
            for (int j = 0; j < ks; j++) {
                P2.scol[j] = P.scol[j];
                P2.dcol[j] = P.dcol[j];
                P2.swidth[j] = P.swidth[j];
            }
            P2.state = P.state; P2.ticket = P.ticket; P2.total = P.total;
        }
        P.num_tiles = tiles_this;
        if (smem > 40 * 1024) { // static shared memory counts against the 48 KB default too
            e = cudaFuncSetAttribute(kern_ptr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) break;
        }
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern_ptr, FT, smem);
        if (e != cudaSuccess) break;
        const int64_t want_occ = ctx->opt("filter.ctas_per_sm", 0);
        if (want_occ > 0) occ = (int)std::min<int64_t>(occ, want_occ);
        occ = std::max(occ, 1);
        const unsigned grid = (unsigned)std::min<int64_t>(tiles_this, (int64_t)ctx->num_sms * occ);
        if (j0 == 0) ctx->kernel_begin();
        if (impl == 1) kern1<<<grid, FT, smem, ctx->stream>>>(P);
        else ch.kern<<<grid, FT, smem, ctx->stream>>>(P2);
        e = cudaGetLastError();
        ctx->count_launch();
        if (j0 + MAXS >= k) ctx->kernel_end();
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ctx->h_scalars, scratch + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ctx->dfree(scratch);
    if (e != cudaSuccess) {
        hk_table_free(ctx, t);
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
