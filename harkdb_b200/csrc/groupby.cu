// groupby.cu — GROUP BY: radix sort by key (sort.cu) + K4 segmented reduction over the sorted rows.
//
// Reference: futhark/groupby.fut:51-62 — project to [g_col]++s_cols (:52-53), sort rows by column 0 (:54),
// head flags (:55-56), segmented reduce with the per-column operator of type_func (:35-41, :58; codes
// 1 prod, 2 sum, 3 max, 4 min, anything else min; u32 wrap-around).  All five operators are commutative and
// associative, so the result does not depend on the reduction order and integer results are bit-exact.
// Extensions: typed columns, COUNT (5), AVG (6), HAVING.
//
// K4 design (HBM-bound; algorithmic bytes = n·(key + value widths) + G·row):
//   1. hk_seg_count: every warp owns a contiguous RANGE of 4096 sorted rows and counts the segment heads in it;
//   2. a tiny exclusive scan turns the per-range counts into the index of each range's first segment (and G);
//   3. hk_seg_reduce: every warp re-derives the head flags of its range (keys are L1/L2-hot), writes the group
//      keys and head positions, then streams each value column once: 4 rows per lane per 128-bit load, lane-
//      local fold, 5-step segmented warp-shuffle scan, carry in a register from one 128-row group to the
//      next.  A segment that lies inside one range is written with a plain store; only segments crossing a
//      range boundary (<= 2 per range) are combined with atomics into the identity-initialised output.
//   4. hk_seg_finalize: COUNT from head positions, AVG = f64 sum / count, f32 SUM rounding.
#include <algorithm>
#include <limits>
#include <type_traits>
#include <new>
#include <stdexcept>
#include <vector>

#include "dense_agg.cuh"
#include "hark_internal.cuh"
#include "sort.cuh"

namespace {

constexpr int GR_T = 256;
constexpr int GR_WARPS = GR_T / 32;
constexpr int GR_WR = 4096;            // rows per warp range
constexpr int GR_GROUPS = GR_WR / 128; // 128-row groups per range (lane owns 4 consecutive rows of a group)
constexpr int MAX_AGG = 16;

enum AccClass { CLS_U32 = 0, CLS_U64 = 1, CLS_F64ACC = 2, CLS_F32MM = 3, CLS_F64MM = 4 };
enum AccOp { OP_PROD = 0, OP_SUM = 1, OP_MAXU = 2, OP_MINU = 3, OP_MAXS = 4, OP_MINS = 5 };

struct AggSpec {
    int cls;      // AccClass
    int op;       // AccOp (MAXS/MINS double as float max/min)
    int in_dtype; // hark_dtype of the input array
    const void *in;
    void *acc;    // [G] accumulator array typed by cls
};

struct SegParams {
    const void *keys;
    int64_t n;
    int64_t num_ranges;
    const unsigned long long *range_base; // [num_ranges] index of the first segment that has its head in the range
    void *out_key;                        // [G] raw key bits
    unsigned long long *out_pos;          // [G] row of each head
    int nagg;
    AggSpec agg[MAX_AGG];
};

template <int KW> struct KRaw;
template <> struct KRaw<4> { using T = uint32_t; };
template <> struct KRaw<8> { using T = uint64_t; };

// 4 consecutive elements starting at row r (r % 4 == 0; the column is padded, see hark_ctx::dalloc)
template <typename T>
__device__ __forceinline__ void load4(const T *p, int64_t r, T (&x)[4]) {
    if constexpr (sizeof(T) == 4) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p + r);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(p + r);
        const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(p + r + 2);
        x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
    }
}

// head bits of this lane's 4 rows of one group.  prev_last = key of the row before the group (valid if r0 > 0).
template <int KW>
__device__ __forceinline__ uint32_t head_bits(const typename KRaw<KW>::T *keys, int64_t g_row0, int64_t n, int lane,
                                              typename KRaw<KW>::T &prev_last, typename KRaw<KW>::T (&k)[4]) {
    using T = typename KRaw<KW>::T;
    const int64_t r = g_row0 + lane * 4;
    k[0] = k[1] = k[2] = k[3] = (T)0;
    if (r < n) load4<T>(keys, r, k);
    T p = __shfl_up_sync(HK_FULL_MASK, k[3], 1);
    if (lane == 0) p = prev_last;
    uint32_t h = 0;
    if (r < n && (r == 0 || k[0] != p)) h |= 1u;
    if (r + 1 < n && k[1] != k[0]) h |= 2u;
    if (r + 2 < n && k[2] != k[1]) h |= 4u;
    if (r + 3 < n && k[3] != k[2]) h |= 8u;
    prev_last = __shfl_sync(HK_FULL_MASK, k[3], 31);
    return h;
}

template <int KW>
__global__ void __launch_bounds__(GR_T) hk_seg_count_kernel(const void *keys_v, int64_t n, int64_t num_ranges,
                                                             uint32_t *range_counts) {
    using T = typename KRaw<KW>::T;
    const T *keys = reinterpret_cast<const T *>(keys_v);
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * GR_WARPS;
    for (int64_t rid = (int64_t)blockIdx.x * GR_WARPS + (threadIdx.x >> 5); rid < num_ranges; rid += nwarps) {
        const int64_t r0 = rid * GR_WR;
        T prev_last = r0 > 0 ? keys[r0 - 1] : (T)0;
        uint32_t cnt = 0;
        for (int g = 0; g < GR_GROUPS; g++) {
            const int64_t g_row0 = r0 + (int64_t)g * 128;
            if (g_row0 >= n) break;
            T k[4];
            cnt += __popc(head_bits<KW>(keys, g_row0, n, lane, prev_last, k));
        }
        cnt = hk_warp_sum_u32(cnt);
        if (lane == 0) range_counts[rid] = cnt;
    }
}

// exclusive scan of u32 counts -> u64 bases, total in base[count]; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) hk_scan_counts_kernel(const uint32_t *counts, unsigned long long *base, int64_t count) {
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
        base[count] = run;
    }
    __syncthreads();
    unsigned long long run = s_part[t];
    for (int64_t i = b; i < e; i++) {
        base[i] = run;
        run += counts[i];
    }
}

// ---- operators ----
template <typename A, int OP> struct OpFn;
template <typename A> struct OpFn<A, OP_PROD> { __device__ __forceinline__ static A f(A a, A b) { return a * b; } };
template <typename A> struct OpFn<A, OP_SUM> { __device__ __forceinline__ static A f(A a, A b) { return a + b; } };
template <> struct OpFn<uint32_t, OP_MAXU> { __device__ __forceinline__ static uint32_t f(uint32_t a, uint32_t b) { return a > b ? a : b; } };
template <> struct OpFn<uint32_t, OP_MINU> { __device__ __forceinline__ static uint32_t f(uint32_t a, uint32_t b) { return a < b ? a : b; } };
template <> struct OpFn<uint32_t, OP_MAXS> { __device__ __forceinline__ static uint32_t f(uint32_t a, uint32_t b) { return (int32_t)a > (int32_t)b ? a : b; } };
template <> struct OpFn<uint32_t, OP_MINS> { __device__ __forceinline__ static uint32_t f(uint32_t a, uint32_t b) { return (int32_t)a < (int32_t)b ? a : b; } };
template <> struct OpFn<uint64_t, OP_MAXU> { __device__ __forceinline__ static uint64_t f(uint64_t a, uint64_t b) { return a > b ? a : b; } };
template <> struct OpFn<uint64_t, OP_MINU> { __device__ __forceinline__ static uint64_t f(uint64_t a, uint64_t b) { return a < b ? a : b; } };
template <> struct OpFn<uint64_t, OP_MAXS> { __device__ __forceinline__ static uint64_t f(uint64_t a, uint64_t b) { return (int64_t)a > (int64_t)b ? a : b; } };
template <> struct OpFn<uint64_t, OP_MINS> { __device__ __forceinline__ static uint64_t f(uint64_t a, uint64_t b) { return (int64_t)a < (int64_t)b ? a : b; } };
template <> struct OpFn<float, OP_MAXS> { __device__ __forceinline__ static float f(float a, float b) { return fmaxf(a, b); } };
template <> struct OpFn<float, OP_MINS> { __device__ __forceinline__ static float f(float a, float b) { return fminf(a, b); } };
template <> struct OpFn<double, OP_MAXS> { __device__ __forceinline__ static double f(double a, double b) { return fmax(a, b); } };
template <> struct OpFn<double, OP_MINS> { __device__ __forceinline__ static double f(double a, double b) { return fmin(a, b); } };

template <typename A, int OP>
__device__ __forceinline__ void atomic_combine(A *addr, A v) {
    if constexpr (sizeof(A) == 4) {
        unsigned int *a = reinterpret_cast<unsigned int *>(addr);
        unsigned int old = *a, assumed;
        do {
            assumed = old;
            A cur;
            memcpy(&cur, &assumed, 4);
            const A nv = OpFn<A, OP>::f(cur, v);
            unsigned int nb;
            memcpy(&nb, &nv, 4);
            old = atomicCAS(a, assumed, nb);
        } while (old != assumed);
    } else {
        unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
        unsigned long long old = *a, assumed;
        do {
            assumed = old;
            A cur;
            memcpy(&cur, &assumed, 8);
            const A nv = OpFn<A, OP>::f(cur, v);
            unsigned long long nb;
            memcpy(&nb, &nv, 8);
            old = atomicCAS(a, assumed, nb);
        } while (old != assumed);
    }
}

// value loaders: 4 rows of the input column as accumulator type A; rows >= n read as `ident`
template <typename A>
__device__ __forceinline__ void load_vals(const void *in, int in_dtype, int64_t r, int64_t n, A ident, A (&x)[4]) {
    x[0] = x[1] = x[2] = x[3] = ident;
    if (r >= n) return;
    if constexpr (sizeof(A) == 4 && !std::is_same<A, float>::value) { // CLS_U32
        uint32_t v[4];
        load4<uint32_t>(reinterpret_cast<const uint32_t *>(in), r, v);
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (r + e < n) x[e] = v[e];
    } else if constexpr (std::is_same<A, float>::value) { // CLS_F32MM
        uint32_t v[4];
        load4<uint32_t>(reinterpret_cast<const uint32_t *>(in), r, v);
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (r + e < n) x[e] = __uint_as_float(v[e]);
    } else if constexpr (std::is_same<A, uint64_t>::value) { // CLS_U64 (4-byte inputs are widened: SUM64)
        if (in_dtype == HARK_I32 || in_dtype == HARK_U32) {
            uint32_t v[4];
            load4<uint32_t>(reinterpret_cast<const uint32_t *>(in), r, v);
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (r + e < n) x[e] = in_dtype == HARK_I32 ? (uint64_t)(int64_t)(int32_t)v[e] : (uint64_t)v[e];
        } else {
            uint64_t v[4];
            load4<uint64_t>(reinterpret_cast<const uint64_t *>(in), r, v);
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (r + e < n) x[e] = v[e];
        }
    } else { // double: CLS_F64ACC (any input dtype) and CLS_F64MM (f64 input)
        if (in_dtype == HARK_F64 || in_dtype == HARK_I64) {
            uint64_t v[4];
            load4<uint64_t>(reinterpret_cast<const uint64_t *>(in), r, v);
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (r + e < n) x[e] = in_dtype == HARK_F64 ? __longlong_as_double((long long)v[e]) : (double)(long long)v[e];
        } else {
            uint32_t v[4];
            load4<uint32_t>(reinterpret_cast<const uint32_t *>(in), r, v);
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (r + e < n)
                    x[e] = in_dtype == HARK_F32 ? (double)__uint_as_float(v[e])
                         : in_dtype == HARK_I32 ? (double)(int32_t)v[e] : (double)v[e];
        }
    }
}

// One value column over one warp range.  hmask_s: this warp's head nibbles in shared memory
// ([GR_GROUPS/8][32] words: word g/8 of lane l holds the nibbles of groups 8*(g/8) .. +7).
template <typename A, int OP>
__device__ __forceinline__ void reduce_range(const AggSpec &ag, A ident, int64_t r0, int64_t n, int lane,
                                             unsigned long long seg_base, const uint32_t *hmask_s) {
    A *out = reinterpret_cast<A *>(ag.acc);
    A carry = ident;
    int carry_started = 0;
    unsigned long long heads_before = seg_base; // heads in rows before the current group (global)
    for (int g = 0; g < GR_GROUPS; g++) {
        const int64_t g_row0 = r0 + (int64_t)g * 128;
        if (g_row0 >= n) break;
        const uint32_t h = (hmask_s[(g >> 3) * 32 + lane] >> ((g & 7) * 4)) & 0xfu;
        A x[4];
        load_vals<A>(ag.in, ag.in_dtype, g_row0 + lane * 4, n, ident, x);
        const uint32_t nh = __popc(h);
        const uint32_t inc = hk_warp_incl_scan_u32(nh);
        const unsigned long long hb = heads_before + (inc - nh); // heads before this lane's rows
        // lane-local fold: `left` = rows before the first head, complete runs are stored, `cur` = open right piece
        A left = ident, cur = ident;
        int seen = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if ((h >> e) & 1u) {
                if (seen == 0) left = cur;
                else out[hb + seen - 1] = cur; // run closed inside this lane
                cur = x[e];
                seen++;
            } else {
                cur = OpFn<A, OP>::f(cur, x[e]);
            }
        }
        // inclusive segmented scan across lanes of (flag = lane has a head, value = open piece)
        A s = cur;
        int sf = seen > 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const A t = __shfl_up_sync(HK_FULL_MASK, s, o);
            const int tf = __shfl_up_sync(HK_FULL_MASK, sf, o);
            if (lane >= o) {
                if (!sf) s = OpFn<A, OP>::f(t, s);
                sf |= tf;
            }
        }
        A ps = __shfl_up_sync(HK_FULL_MASK, s, 1);
        int psf = __shfl_up_sync(HK_FULL_MASK, sf, 1);
        if (lane == 0) {
            ps = ident;
            psf = 0;
        }
        if (seen > 0 && hb > 0) { // this lane's first head closes segment hb-1
            const A tprev = psf ? ps : OpFn<A, OP>::f(carry, ps);
            const A tot = OpFn<A, OP>::f(tprev, left);
            if (psf | carry_started) out[hb - 1] = tot;       // the segment began inside this range
            else atomic_combine<A, OP>(&out[hb - 1], tot);    // it began in an earlier range
        }
        const A s31 = __shfl_sync(HK_FULL_MASK, s, 31);
        const int f31 = __shfl_sync(HK_FULL_MASK, sf, 31);
        carry = f31 ? s31 : OpFn<A, OP>::f(carry, s31);
        carry_started |= f31;
        heads_before += __shfl_sync(HK_FULL_MASK, inc, 31);
    }
    // the segment still open at the end of the range may continue in the next range
    if (lane == 0 && heads_before > 0) atomic_combine<A, OP>(&out[heads_before - 1], carry);
}

template <typename A>
__device__ __forceinline__ void reduce_dispatch_int(const AggSpec &ag, int64_t r0, int64_t n, int lane,
                                                    unsigned long long seg_base, const uint32_t *hm) {
    constexpr A smin = (A)1 << (sizeof(A) * 8 - 1);
    switch (ag.op) {
    case OP_PROD: reduce_range<A, OP_PROD>(ag, (A)1, r0, n, lane, seg_base, hm); break;
    case OP_SUM: reduce_range<A, OP_SUM>(ag, (A)0, r0, n, lane, seg_base, hm); break;
    case OP_MAXU: reduce_range<A, OP_MAXU>(ag, (A)0, r0, n, lane, seg_base, hm); break;
    case OP_MINU: reduce_range<A, OP_MINU>(ag, (A) ~(A)0, r0, n, lane, seg_base, hm); break;
    case OP_MAXS: reduce_range<A, OP_MAXS>(ag, smin, r0, n, lane, seg_base, hm); break;
    default: reduce_range<A, OP_MINS>(ag, (A)(smin - 1), r0, n, lane, seg_base, hm); break;
    }
}

template <int KW>
__global__ void __launch_bounds__(GR_T) hk_seg_reduce_kernel(const __grid_constant__ SegParams P) {
    using T = typename KRaw<KW>::T;
    __shared__ uint32_t s_hmask[GR_WARPS][(GR_GROUPS / 8) * 32];
    const T *keys = reinterpret_cast<const T *>(P.keys);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *hm = s_hmask[warp];
    const int64_t nwarps = (int64_t)gridDim.x * GR_WARPS;
    for (int64_t rid = (int64_t)blockIdx.x * GR_WARPS + warp; rid < P.num_ranges; rid += nwarps) {
        const int64_t r0 = rid * GR_WR;
        const unsigned long long seg_base = P.range_base[rid];
        // ---- pass 0: head flags -> shared memory; group keys and head rows -> output ----
        T prev_last = r0 > 0 ? keys[r0 - 1] : (T)0;
        unsigned long long heads_before = seg_base;
        uint32_t word = 0;
        for (int g = 0; g < GR_GROUPS; g++) {
            const int64_t g_row0 = r0 + (int64_t)g * 128;
            uint32_t h = 0;
            if (g_row0 < P.n) { // warp-uniform
                T k[4];
                h = head_bits<KW>(keys, g_row0, P.n, lane, prev_last, k);
                const uint32_t nh = __popc(h);
                const uint32_t inc = hk_warp_incl_scan_u32(nh);
                unsigned long long seg = heads_before + (inc - nh);
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if ((h >> e) & 1u) {
                        reinterpret_cast<T *>(P.out_key)[seg] = k[e];
                        P.out_pos[seg] = (unsigned long long)(g_row0 + lane * 4 + e);
                        seg++;
                    }
                heads_before += __shfl_sync(HK_FULL_MASK, inc, 31);
            }
            word |= h << ((g & 7) * 4);
            if ((g & 7) == 7) {
                hm[(g >> 3) * 32 + lane] = word;
                word = 0;
            }
        }
        __syncwarp();
        // ---- one streaming pass per aggregate ----
        for (int a = 0; a < P.nagg; a++) {
            const AggSpec &ag = P.agg[a];
            switch (ag.cls) {
            case CLS_U32: reduce_dispatch_int<uint32_t>(ag, r0, P.n, lane, seg_base, hm); break;
            case CLS_U64: reduce_dispatch_int<uint64_t>(ag, r0, P.n, lane, seg_base, hm); break;
            case CLS_F64ACC:
                if (ag.op == OP_PROD) reduce_range<double, OP_PROD>(ag, 1.0, r0, P.n, lane, seg_base, hm);
                else reduce_range<double, OP_SUM>(ag, 0.0, r0, P.n, lane, seg_base, hm);
                break;
            case CLS_F32MM:
                if (ag.op == OP_MAXS) reduce_range<float, OP_MAXS>(ag, -INFINITY, r0, P.n, lane, seg_base, hm);
                else reduce_range<float, OP_MINS>(ag, INFINITY, r0, P.n, lane, seg_base, hm);
                break;
            default:
                if (ag.op == OP_MAXS) reduce_range<double, OP_MAXS>(ag, -(double)INFINITY, r0, P.n, lane, seg_base, hm);
                else reduce_range<double, OP_MINS>(ag, (double)INFINITY, r0, P.n, lane, seg_base, hm);
                break;
            }
        }
        __syncwarp();
    }
}

// identity fill of an accumulator array
template <typename A>
__global__ void __launch_bounds__(256) hk_fill_kernel(A *p, A v, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

enum FinKind { FIN_NONE = 0, FIN_COUNT = 1, FIN_AVG = 2, FIN_F32 = 3 };
struct FinParams {
    int64_t G, n;
    const unsigned long long *pos;
    int nout;
    int kind[MAX_AGG];
    void *acc[MAX_AGG]; // f64 accumulator (AVG in place, F32 source)
    void *dst[MAX_AGG]; // COUNT: i64 column; F32: f32 column
};

__global__ void __launch_bounds__(256) hk_seg_finalize_kernel(const __grid_constant__ FinParams P) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < P.G; g += stride) {
        const unsigned long long cnt = (g + 1 < P.G ? P.pos[g + 1] : (unsigned long long)P.n) - P.pos[g];
        for (int j = 0; j < P.nout; j++) {
            switch (P.kind[j]) {
            case FIN_COUNT: reinterpret_cast<long long *>(P.dst[j])[g] = (long long)cnt; break;
            case FIN_AVG: reinterpret_cast<double *>(P.acc[j])[g] /= (double)cnt; break;
            case FIN_F32: reinterpret_cast<float *>(P.dst[j])[g] = (float)reinterpret_cast<const double *>(P.acc[j])[g]; break;
            default: break;
            }
        }
    }
}

unsigned grid_for(hark_ctx *ctx, int64_t n, int per_sm = 8) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * per_sm));
}

template <typename A>
int fill(hark_ctx *ctx, void *p, A v, int64_t n) {
    if (n == 0) return HARK_OK;
    hk_fill_kernel<A><<<grid_for(ctx, n), 256, 0, ctx->stream>>>((A *)p, v, n);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch();
    return HARK_OK;
}

struct Bufs { // scratch that must be released on every exit path
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
    void release(void *p) { // ownership moves elsewhere
        v.erase(std::remove(v.begin(), v.end(), p), v.end());
    }
};

} // namespace

// Segmented reduction over rows already sorted by key.  key / vals are device arrays in sorted order.
// aggs: (value array index or -1, hark_agg code).  Produces [key, agg_1..agg_c] as a new table.
int hk_segmented_aggregate(hark_ctx *ctx, hark_table **out, int64_t n, const void *sorted_key, int32_t key_dtype,
                           const std::vector<const void *> &vals, const std::vector<int32_t> &val_dtypes,
                           const std::vector<std::pair<int, int>> &aggs, bool pinned_u32) {
    const int c = (int)aggs.size();
    if (c > MAX_AGG) return ctx->fail(HARK_ERR_UNSUPPORTED, "groupby: more than 16 aggregates");
    const int kw = hk_dtype_size(key_dtype);
    Bufs scratch(ctx);

    // output dtypes
    std::vector<int32_t> odt(1 + c);
    odt[0] = pinned_u32 ? HARK_U32 : key_dtype;
    for (int j = 0; j < c; j++) {
        const int code = aggs[j].second;
        const int32_t vdt = aggs[j].first >= 0 ? val_dtypes[aggs[j].first] : HARK_I64;
        if (code == HARK_AGG_COUNT) odt[1 + j] = HARK_I64;
        else if (code == HARK_AGG_AVG || code == HARK_AGG_SUMF64) odt[1 + j] = HARK_F64;
        else if (code == HARK_AGG_SUM64 && !pinned_u32) odt[1 + j] = hk_dtype_int(vdt) ? HARK_I64 : HARK_F64;
        else odt[1 + j] = pinned_u32 ? HARK_U32 : vdt;
    }
    if (n == 0) return hk_table_alloc(ctx, out, 0, 0, odt.data(), 1 + c);

    // ---- 1. heads per range, 2. scan ----
    const int64_t num_ranges = (n + GR_WR - 1) / GR_WR;
    uint32_t *d_counts = nullptr;
    unsigned long long *d_base = nullptr;
    HK_TRY(scratch.alloc((void **)&d_counts, sizeof(uint32_t) * (size_t)num_ranges));
    HK_TRY(scratch.alloc((void **)&d_base, sizeof(unsigned long long) * (size_t)(num_ranges + 1)));
    const unsigned wgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((num_ranges + GR_WARPS - 1) / GR_WARPS, (int64_t)ctx->num_sms * 8));
    if (kw == 4) hk_seg_count_kernel<4><<<wgrid, GR_T, 0, ctx->stream>>>(sorted_key, n, num_ranges, d_counts);
    else hk_seg_count_kernel<8><<<wgrid, GR_T, 0, ctx->stream>>>(sorted_key, n, num_ranges, d_counts);
    HK_CHECK_LAUNCH(ctx);
    hk_scan_counts_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, d_base, num_ranges);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch(2);
    HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, d_base + num_ranges, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t G = (int64_t)ctx->h_scalars[0];

    // ---- output table + accumulators ----
    hark_table *t = nullptr;
    HK_TRY(hk_table_alloc(ctx, &t, G, G, odt.data(), 1 + c));
    struct Guard {
        hark_ctx *ctx;
        hark_table *t;
        ~Guard() { if (t) hk_table_free(ctx, t); }
    } guard{ctx, t};
    unsigned long long *d_pos = nullptr;
    HK_TRY(scratch.alloc((void **)&d_pos, sizeof(unsigned long long) * (size_t)G));

    SegParams P;
    memset(&P, 0, sizeof P);
    P.keys = sorted_key;
    P.n = n;
    P.num_ranges = num_ranges;
    P.range_base = d_base;
    P.out_key = t->cols[0].ptr;
    P.out_pos = d_pos;
    FinParams F;
    memset(&F, 0, sizeof F);
    F.G = G;
    F.n = n;
    F.pos = d_pos;
    F.nout = c;
    int nagg = 0;
    for (int j = 0; j < c; j++) {
        int code = aggs[j].second;
        if (code == HARK_AGG_COUNT) {
            F.kind[j] = FIN_COUNT;
            F.dst[j] = t->cols[1 + j].ptr;
            continue;
        }
        if (code < HARK_AGG_PROD || code > HARK_AGG_SUM64 || (pinned_u32 && code > HARK_AGG_MIN)) code = HARK_AGG_MIN; // groupby.fut:41
        const int vi = aggs[j].first;
        const int32_t vdt = pinned_u32 ? HARK_U32 : val_dtypes[vi];
        AggSpec &ag = P.agg[nagg++];
        ag.in = vals[vi];
        ag.in_dtype = vdt;
        const bool is_f = (vdt == HARK_F32 || vdt == HARK_F64);
        const bool is_signed = (vdt == HARK_I32 || vdt == HARK_I64);
        const int w = hk_dtype_size(vdt);
        if (code == HARK_AGG_SUM64 && !is_f) { // exact integer sum: 64-bit accumulator over sign- / zero-extended inputs
            ag.cls = CLS_U64;
            ag.op = OP_SUM;
            ag.acc = t->cols[1 + j].ptr;
            HK_TRY(fill<uint64_t>(ctx, ag.acc, 0ull, G));
            continue;
        }
        if (code == HARK_AGG_SUM64) code = HARK_AGG_SUMF64;
        if (code == HARK_AGG_AVG || code == HARK_AGG_SUMF64 || (is_f && (code == HARK_AGG_SUM || code == HARK_AGG_PROD))) {
            ag.cls = CLS_F64ACC;
            ag.op = code == HARK_AGG_PROD ? OP_PROD : OP_SUM;
            if (code == HARK_AGG_AVG || code == HARK_AGG_SUMF64 || vdt == HARK_F64) {
                ag.acc = t->cols[1 + j].ptr; // f64 output column doubles as the accumulator
                if (code == HARK_AGG_AVG) {
                    F.kind[j] = FIN_AVG;
                    F.acc[j] = ag.acc;
                }
            } else { // f32 SUM / PROD: accumulate in f64, round once at the end
                HK_TRY(scratch.alloc(&ag.acc, sizeof(double) * (size_t)std::max<int64_t>(G, 1)));
                F.kind[j] = FIN_F32;
                F.acc[j] = ag.acc;
                F.dst[j] = t->cols[1 + j].ptr;
            }
            HK_TRY(fill<double>(ctx, ag.acc, ag.op == OP_PROD ? 1.0 : 0.0, G));
        } else if (is_f) {
            ag.cls = w == 4 ? CLS_F32MM : CLS_F64MM;
            ag.op = code == HARK_AGG_MAX ? OP_MAXS : OP_MINS;
            ag.acc = t->cols[1 + j].ptr;
            if (w == 4) HK_TRY(fill<float>(ctx, ag.acc, code == HARK_AGG_MAX ? -INFINITY : INFINITY, G));
            else HK_TRY(fill<double>(ctx, ag.acc, code == HARK_AGG_MAX ? -(double)INFINITY : (double)INFINITY, G));
        } else {
            ag.cls = w == 4 ? CLS_U32 : CLS_U64;
            ag.op = code == HARK_AGG_PROD ? OP_PROD : code == HARK_AGG_SUM ? OP_SUM
                  : code == HARK_AGG_MAX ? (is_signed ? OP_MAXS : OP_MAXU) : (is_signed ? OP_MINS : OP_MINU);
            ag.acc = t->cols[1 + j].ptr;
            if (w == 4) {
                const uint32_t id = ag.op == OP_PROD ? 1u : ag.op == OP_SUM ? 0u : ag.op == OP_MAXU ? 0u
                                  : ag.op == OP_MINU ? 0xffffffffu : ag.op == OP_MAXS ? 0x80000000u : 0x7fffffffu;
                HK_TRY(fill<uint32_t>(ctx, ag.acc, id, G));
            } else {
                const uint64_t id = ag.op == OP_PROD ? 1ull : ag.op == OP_SUM ? 0ull : ag.op == OP_MAXU ? 0ull
                                  : ag.op == OP_MINU ? ~0ull : ag.op == OP_MAXS ? 0x8000000000000000ull : 0x7fffffffffffffffull;
                HK_TRY(fill<uint64_t>(ctx, ag.acc, id, G));
            }
        }
    }
    P.nagg = nagg;

    // ---- 3. reduce, 4. finalize ----
    ctx->kernel_begin();
    if (kw == 4) hk_seg_reduce_kernel<4><<<wgrid, GR_T, 0, ctx->stream>>>(P);
    else hk_seg_reduce_kernel<8><<<wgrid, GR_T, 0, ctx->stream>>>(P);
    HK_CHECK_LAUNCH(ctx);
    ctx->kernel_end();
    hk_seg_finalize_kernel<<<grid_for(ctx, G), 256, 0, ctx->stream>>>(F);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch(2);
    guard.t = nullptr;
    *out = t;
    return HARK_OK;
}

int hk_groupby(hark_ctx *ctx, hark_table **out, const hark_table *db, int32_t g_col, const int32_t *s_cols,
               const int32_t *ops, int64_t c, const hark_pred *having, int64_t nh, bool pinned_u32) {
    const int64_t n = db->n, m = (int64_t)db->cols.size();
    // groupby.fut:52 indexes every row with g_col and s_cols: out of bounds is an error (only when rows exist)
    if (n > 0 || !pinned_u32) {
        HK_ARG(ctx, g_col >= 0 && g_col < m, "query_groupby: group column index out of bounds");
        for (int64_t j = 0; j < c; j++)
            HK_ARG(ctx, s_cols[j] >= 0 && s_cols[j] < m, "query_groupby: aggregated column index out of bounds");
    }
    if (pinned_u32) {
        for (int64_t col = 0; col < m; col++)
            HK_ARG(ctx, db->cols[col].dtype == HARK_I32 || db->cols[col].dtype == HARK_U32,
                   "query_groupby: the reference entry takes u32 data (use hark_entry_query_groupby_ex for typed tables)");
    } else {
        HK_ARG(ctx, m > 0 && hk_dtype_int(db->cols[g_col].dtype), "query_groupby_ex: the group key must be an integer column");
    }
    if (pinned_u32 && n == 0) { // groupby.fut with zero rows: [0][1+c]
        std::vector<int32_t> odt((size_t)(1 + c), HARK_U32);
        ctx->entry_begin();
        HK_TRY(hk_table_alloc(ctx, out, 0, 0, odt.data(), 1 + c));
        ctx->entry_end(0, 0, 0);
        return HARK_OK;
    }
    ctx->entry_begin();
    const int32_t key_dtype = pinned_u32 ? HARK_U32 : db->cols[g_col].dtype;

    // ---- K2: dense / partitioned shared-memory aggregation when the key range allows it (dense_agg.cu) ----
    if (ctx->opt("groupby.impl", 0) != 1 && n > 0) {
        hk_dense_req rq;
        rq.n = n;
        rq.key = db->cols[g_col].ptr;
        rq.key_dtype = key_dtype;
        rq.out_key_dtype = key_dtype;
        rq.pinned_u32 = pinned_u32;
        rq.c = (int)c;
        bool eligible = c <= HK_DENSE_MAX_AGGS && hk_dtype_size(db->cols[g_col].dtype) == hk_dtype_size(key_dtype);
        std::vector<int> val_of_col((size_t)m, -1);
        for (int64_t j = 0; j < c && eligible; j++) {
            int code = ops[j];
            if (pinned_u32 && (code < HARK_AGG_PROD || code > HARK_AGG_MIN)) code = HARK_AGG_MIN; // groupby.fut:41
            if (!pinned_u32 && (code < HARK_AGG_PROD || code > HARK_AGG_SUM64)) code = HARK_AGG_MIN;
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
                rq.vals[rq.nvals] = db->cols[col].ptr;
                rq.val_cols[rq.nvals] = &db->cols[col];
                rq.val_dtypes[rq.nvals] = pinned_u32 ? HARK_U32 : db->cols[col].dtype;
                rq.nvals++;
            }
            rq.agg_val[j] = val_of_col[col];
        }
        if (eligible) {
            HK_TRY(hk_column_minmax(ctx, db->cols[g_col], n, key_dtype, &rq.g_lo, &rq.g_hi));
            bool handled = false;
            hark_table *t = nullptr;
            HK_TRY(hk_dense_groupby(ctx, &t, rq, &handled));
            if (handled) {
                if (nh > 0) { // HAVING: K1 over the (small) group table; output column indices
                    std::vector<int32_t> all;
                    for (int64_t j = 0; j < 1 + c; j++) all.push_back((int32_t)j);
                    hark_table *f = nullptr;
                    int rc = hk_filter(ctx, &f, t, all.data(), 1 + c, having, nh);
                    hk_table_free(ctx, t);
                    if (rc != HARK_OK) return rc;
                    t = f;
                }
                int64_t alg = n * hk_dtype_size(key_dtype);
                for (int v = 0; v < rq.nvals; v++) alg += n * 4;
                for (auto &col : t->cols) alg += t->n * hk_dtype_size(col.dtype);
                ctx->entry_end(alg, n, t->n);
                *out = t;
                return HARK_OK;
            }
        }
    }

    // ---- sort path: carried arrays = the key column first, then each distinct value column ----
    std::vector<hk_sort_array> arrays;
    std::vector<int> array_of_col((size_t)m, -1);
    {
        hk_sort_array a;
        a.in = db->cols[g_col].ptr;
        a.width = hk_dtype_size(db->cols[g_col].dtype);
        arrays.push_back(a);
        array_of_col[g_col] = 0;
    }
    for (int64_t j = 0; j < c; j++) {
        if (!pinned_u32 && ops[j] == HARK_AGG_COUNT) continue; // COUNT needs no value column
        const int col = s_cols[j];
        if (array_of_col[col] >= 0) continue;
        hk_sort_array a;
        a.in = db->cols[col].ptr;
        a.width = hk_dtype_size(db->cols[col].dtype);
        array_of_col[col] = (int)arrays.size();
        arrays.push_back(a);
    }
    if ((int)arrays.size() > HK_SORT_MAX_ARRAYS)
        return ctx->fail(HARK_ERR_UNSUPPORTED, "query_groupby: more than 19 distinct aggregated columns");
    std::vector<hk_sort_keyspec> keys{hk_sort_keyspec{0, key_dtype, 0}};
    hk_sort_info info;
    HK_TRY(hk_radix_sort(ctx, n, keys, arrays, 0, nullptr, &info));

    std::vector<const void *> vals;
    std::vector<int32_t> vdts;
    std::vector<int> val_of_array(arrays.size(), -1);
    std::vector<std::pair<int, int>> aggs;
    for (int64_t j = 0; j < c; j++) {
        int code = ops[j];
        if (pinned_u32 && (code < HARK_AGG_PROD || code > HARK_AGG_MIN)) code = HARK_AGG_MIN; // groupby.fut:41
        if (!pinned_u32 && code == HARK_AGG_COUNT) {
            aggs.push_back({-1, code});
            continue;
        }
        const int ai = array_of_col[s_cols[j]];
        if (val_of_array[ai] < 0) {
            val_of_array[ai] = (int)vals.size();
            vals.push_back(arrays[ai].result);
            vdts.push_back(db->cols[s_cols[j]].dtype);
        }
        aggs.push_back({val_of_array[ai], code});
    }
    hark_table *t = nullptr;
    int rc = hk_segmented_aggregate(ctx, &t, n, arrays[0].result, key_dtype, vals, vdts, aggs, pinned_u32);
    for (auto &a : arrays) ctx->dfree(a.result);
    if (rc != HARK_OK) return rc;

    if (nh > 0) { // HAVING: K1 over the (small) group table; output column indices
        std::vector<int32_t> all;
        for (int64_t j = 0; j < 1 + c; j++) all.push_back((int32_t)j);
        hark_table *f = nullptr;
        rc = hk_filter(ctx, &f, t, all.data(), 1 + c, having, nh);
        hk_table_free(ctx, t);
        if (rc != HARK_OK) return rc;
        t = f;
    }
    int64_t alg = 0;
    for (auto &a : arrays) alg += n * a.width;
    for (auto &col : t->cols) alg += t->n * hk_dtype_size(col.dtype);
    ctx->entry_end(alg, n, t->n);
    *out = t;
    return HARK_OK;
}
