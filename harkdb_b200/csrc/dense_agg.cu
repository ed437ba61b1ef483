// dense_agg.cu — K2: GROUP BY without a sort, for keys whose range is small enough to address directly.
//
// Reference: futhark/groupby.fut:51-62 sorts whole rows with 32 one-bit radix passes (:8-22) just to bring equal
// keys together, then runs a segmented scan (:58).  Every operator of type_func (:35-41) is commutative and
// associative, so the grouping can be done by ADDRESS instead of by order, and integer results stay bit-exact.
//
//   R = max(key) - min(key) + 1 "dense slots".  Every CTA keeps a table of K slots in shared memory
//   (K = largest power of two whose accumulators fit ~200 KB: 16384 slots for SUM+COUNT+AVG of one column).
//     R <= K        : one streaming pass; rows are aggregated with shared-memory atomics, tables are merged into
//                     dense global accumulators with one global atomic per (touched slot, accumulator);
//     R <= 256 * K  : one partition pass by (key - min) >> log2(K) first (hk_part_kernel, below), so that every
//                     bucket's key sub-range fits the table; CTAs then walk bucket-contiguous chunks;
//     otherwise     : not handled here (the caller sorts: sort.cu + groupby.cu).
//   A final compaction turns the dense accumulators into the output table — slot order IS ascending key order,
//   so the result needs no sort either.
// join + GROUP BY (config 5) is the same kernel with the slot taken from a lookup: slot+1 = lut[fk - pk_min].
//
// HBM traffic (4-byte key + one 4-byte value, n rows): streaming pass 8n; with the partition pass 4n (histogram)
// + 16n (move) + 8n = 28n, against 2·3·8n + ... for the 3-pass sort path.
#include <algorithm>
#include <new>
#include <stdexcept>
#include <vector>

#include "dense_agg.cuh"
#include "hark_internal.cuh"

namespace {

template <int KW> struct KRaw;
template <> struct KRaw<4> { using T = uint32_t; };
template <> struct KRaw<8> { using T = uint64_t; };

template <int KW>
__device__ __forceinline__ uint64_t ordkey_of(typename KRaw<KW>::T raw, int dtype) {
    if constexpr (KW == 4) return (uint64_t)hk_ordkey32(raw, dtype);
    else return hk_ordkey64(raw, dtype);
}

// ------------------------------------------------------------------------------------------------
// min / max order key
// ------------------------------------------------------------------------------------------------
template <int KW>
__global__ void __launch_bounds__(256) hk_dminmax_kernel(const void *__restrict__ col, int64_t n, int dtype,
                                                          unsigned long long *out /* [0]=min [1]=max */) {
    using T = typename KRaw<KW>::T;
    const T *p = reinterpret_cast<const T *>(col);
    uint64_t lo = ~0ull, hi = 0;
    constexpr int V = 16 / KW; // elements per 128-bit load
    const int64_t nvec = n / V;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        T x[V];
        if constexpr (KW == 4) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(p) + i);
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        } else {
            const ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2 *>(p) + i);
            x[0] = v.x; x[1] = v.y;
        }
#pragma unroll
        for (int e = 0; e < V; e++) {
            const uint64_t u = ordkey_of<KW>(x[e], dtype);
            lo = min(lo, u);
            hi = max(hi, u);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - nvec * V)) {
        const uint64_t u = ordkey_of<KW>(p[nvec * V + threadIdx.x], dtype);
        lo = min(lo, u);
        hi = max(hi, u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(HK_FULL_MASK, lo, o));
        hi = max(hi, __shfl_xor_sync(HK_FULL_MASK, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, (unsigned long long)lo);
        atomicMax(out + 1, (unsigned long long)hi);
    }
}

constexpr int PMAXV = 3; // value arrays the partition pass can carry (partition.cu)

// ------------------------------------------------------------------------------------------------
// K2 aggregation kernel
// ------------------------------------------------------------------------------------------------
// threads per CTA (one CTA per SM: the table takes most of the shared memory)
__host__ __device__ constexpr int dagg_threads(int nv) { return nv <= 1 ? 1024 : 512; }
constexpr int AMAXACC = 12;

enum AccKind {
    A_SUM32 = 0,  // u32 wrap-around sum                      (1 word;  global u32)
    A_SUM64S,     // exact sum of sign-extended i32           (2 words; global u64)
    A_SUM64U,     // exact sum of zero-extended u32           (2 words; global u64)
    A_FSUM,       // f32 -> f64 sum                           (u64 word pair; global f64)
    A_FPROD,      // f32 -> f64 product
    A_MINU, A_MAXU, A_MINS, A_MAXS,   // 32-bit integer min / max
    A_MINF, A_MAXF,                   // f32 min / max (fminf / fmaxf: NaN ignored, like the sort path)
    A_PROD32      // u32 wrap-around product
};

struct DAcc {
    int vcol;
    int kind;
    int word;   // first table word (units of K u32)
    void *gacc; // dense global accumulator [R]
};

struct DAggParams {
    const void *key;
    int key_dtype;
    int64_t n;
    uint64_t g_lo;
    int shift;     // log2(K)
    uint32_t K;
    int nwords;    // u32 words per slot
    int nbins;
    const unsigned long long *offsets; // [nbins + 1] row offsets of the buckets
    const uint32_t *lut;
    long long pk_min, pk_span;
    int nvals;
    const uint32_t *vals[HK_DENSE_MAX_VALS];
    int nacc;
    DAcc acc[AMAXACC];
    unsigned long long *gcnt; // [R]
    int64_t chunk;            // rows per chunk (multiple of 4)
};

__device__ __forceinline__ uint32_t acc_identity(int kind, int w /* 0 or 1 for two-word kinds */) {
    switch (kind) {
    case A_MINU: return 0xffffffffu;
    case A_MINS: return 0x7fffffffu;
    case A_MAXS: return 0x80000000u;
    case A_MINF: return 0x7f800000u;  // +inf
    case A_MAXF: return 0xff800000u;  // -inf
    case A_PROD32: return 1u;
    case A_FPROD: return w == 0 ? 0u : 0x3ff00000u; // 1.0 (little endian: low word first)
    default: return 0u;
    }
}

__device__ __forceinline__ void smem_f64_update(unsigned long long *a, double v, bool mul) {
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        const double cur = __longlong_as_double((long long)assumed);
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(mul ? cur * v : cur + v));
    } while (old != assumed);
}

// one row into the CTA's table (KIND is a compile-time AccKind: the switch over kinds sits outside the row loops)
template <int KIND>
__device__ __forceinline__ void acc_row(uint32_t *tab, uint32_t K, uint32_t idx, int word, uint32_t x) {
    uint32_t *w0 = tab + (size_t)word * K + idx;
    if constexpr (KIND == A_SUM32) {
        atomicAdd(w0, x);
    } else if constexpr (KIND == A_SUM64S) {
        const uint32_t old = atomicAdd(w0, x);
        const int delta = (int)((uint32_t)(old + x) < old) - (int)((int32_t)x < 0);
        if (delta != 0) atomicAdd(reinterpret_cast<int *>(w0 + K), delta);
    } else if constexpr (KIND == A_SUM64U) {
        const uint32_t old = atomicAdd(w0, x);
        if ((uint32_t)(old + x) < old) atomicAdd(w0 + K, 1u);
    } else if constexpr (KIND == A_FSUM || KIND == A_FPROD) {
        smem_f64_update(reinterpret_cast<unsigned long long *>(tab + (size_t)word * K) + idx, (double)__uint_as_float(x), KIND == A_FPROD);
    } else if constexpr (KIND == A_MINU) {
        atomicMin(w0, x);
    } else if constexpr (KIND == A_MAXU) {
        atomicMax(w0, x);
    } else if constexpr (KIND == A_MINS) {
        atomicMin(reinterpret_cast<int *>(w0), (int)x);
    } else if constexpr (KIND == A_MAXS) {
        atomicMax(reinterpret_cast<int *>(w0), (int)x);
    } else if constexpr (KIND == A_MINF || KIND == A_MAXF) {
        uint32_t old = *w0, assumed;
        do {
            assumed = old;
            const float cur = __uint_as_float(assumed);
            const float nv = KIND == A_MINF ? fminf(cur, __uint_as_float(x)) : fmaxf(cur, __uint_as_float(x));
            if (__float_as_uint(nv) == assumed) break;
            old = atomicCAS(w0, assumed, __float_as_uint(nv));
        } while (old != assumed);
    } else { // A_PROD32
        uint32_t old = *w0, assumed;
        do {
            assumed = old;
            old = atomicCAS(w0, assumed, assumed * x);
        } while (old != assumed);
    }
}

template <int KIND, int U>
__device__ __forceinline__ void acc_rows(uint32_t *tab, uint32_t K, const uint32_t (&idx)[U][4], int word, const uint32_t (&x)[U][4]) {
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (idx[u][e] != 0xffffffffu) acc_row<KIND>(tab, K, idx[u][e], word, x[u][e]);
}

template <int U>
__device__ __forceinline__ void acc_dispatch(uint32_t *tab, uint32_t K, const uint32_t (&idx)[U][4], int kind, int word,
                                             const uint32_t (&x)[U][4]) {
    switch (kind) {
    case A_SUM32: acc_rows<A_SUM32, U>(tab, K, idx, word, x); break;
    case A_SUM64S: acc_rows<A_SUM64S, U>(tab, K, idx, word, x); break;
    case A_SUM64U: acc_rows<A_SUM64U, U>(tab, K, idx, word, x); break;
    case A_FSUM: acc_rows<A_FSUM, U>(tab, K, idx, word, x); break;
    case A_FPROD: acc_rows<A_FPROD, U>(tab, K, idx, word, x); break;
    case A_MINU: acc_rows<A_MINU, U>(tab, K, idx, word, x); break;
    case A_MAXU: acc_rows<A_MAXU, U>(tab, K, idx, word, x); break;
    case A_MINS: acc_rows<A_MINS, U>(tab, K, idx, word, x); break;
    case A_MAXS: acc_rows<A_MAXS, U>(tab, K, idx, word, x); break;
    case A_MINF: acc_rows<A_MINF, U>(tab, K, idx, word, x); break;
    case A_MAXF: acc_rows<A_MAXF, U>(tab, K, idx, word, x); break;
    default: acc_rows<A_PROD32, U>(tab, K, idx, word, x); break;
    }
}

// merge one touched slot of the CTA's table into the dense global accumulators, and reset it
__device__ __forceinline__ void flush_slot(uint32_t *tab, uint32_t K, uint32_t i, uint64_t r, const DAggParams &P) {
    const uint32_t c = tab[i];
    if (c == 0) return;
    atomicAdd(&P.gcnt[r], (unsigned long long)c);
    tab[i] = 0;
    for (int ai = 0; ai < P.nacc; ai++) {
        const DAcc &a = P.acc[ai];
        uint32_t *w0 = tab + (size_t)a.word * K + i;
        switch (a.kind) {
        case A_SUM32:
            atomicAdd(reinterpret_cast<uint32_t *>(a.gacc) + r, *w0);
            *w0 = 0;
            break;
        case A_SUM64S:
        case A_SUM64U: {
            const unsigned long long v = ((unsigned long long)(long long)(int32_t)w0[K] << 32) + (unsigned long long)*w0;
            atomicAdd(reinterpret_cast<unsigned long long *>(a.gacc) + r, v);
            *w0 = 0;
            w0[K] = 0;
            break;
        }
        case A_FSUM: {
            unsigned long long *p = reinterpret_cast<unsigned long long *>(tab + (size_t)a.word * K) + i;
            atomicAdd(reinterpret_cast<double *>(a.gacc) + r, __longlong_as_double((long long)*p));
            *p = 0ull;
            break;
        }
        case A_FPROD: {
            unsigned long long *p = reinterpret_cast<unsigned long long *>(tab + (size_t)a.word * K) + i;
            const double v = __longlong_as_double((long long)*p);
            unsigned long long *g = reinterpret_cast<unsigned long long *>(a.gacc) + r;
            unsigned long long old = *g, assumed;
            do {
                assumed = old;
                old = atomicCAS(g, assumed, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)assumed) * v));
            } while (old != assumed);
            *p = 0x3ff0000000000000ull;
            break;
        }
        case A_MINU: atomicMin(reinterpret_cast<uint32_t *>(a.gacc) + r, *w0); *w0 = 0xffffffffu; break;
        case A_MAXU: atomicMax(reinterpret_cast<uint32_t *>(a.gacc) + r, *w0); *w0 = 0u; break;
        case A_MINS: atomicMin(reinterpret_cast<int *>(a.gacc) + r, (int)*w0); *w0 = 0x7fffffffu; break;
        case A_MAXS: atomicMax(reinterpret_cast<int *>(a.gacc) + r, (int)*w0); *w0 = 0x80000000u; break;
        case A_MINF:
        case A_MAXF: {
            uint32_t *g = reinterpret_cast<uint32_t *>(a.gacc) + r;
            const float v = __uint_as_float(*w0);
            uint32_t old = *g, assumed;
            do {
                assumed = old;
                const float cur = __uint_as_float(assumed);
                const float nv = a.kind == A_MINF ? fminf(cur, v) : fmaxf(cur, v);
                if (__float_as_uint(nv) == assumed) break;
                old = atomicCAS(g, assumed, __float_as_uint(nv));
            } while (old != assumed);
            *w0 = acc_identity(a.kind, 0);
            break;
        }
        default: { // A_PROD32
            uint32_t *g = reinterpret_cast<uint32_t *>(a.gacc) + r;
            const uint32_t v = *w0;
            uint32_t old = *g, assumed;
            do {
                assumed = old;
                old = atomicCAS(g, assumed, assumed * v);
            } while (old != assumed);
            *w0 = 1u;
            break;
        }
        }
    }
}

template <int KW>
__device__ __forceinline__ void load_keys4(const typename KRaw<KW>::T *p, int64_t r, typename KRaw<KW>::T (&k)[4]) {
    if constexpr (KW == 4) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(p + r));
        k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
    } else {
        const ulonglong2 a = __ldcs(reinterpret_cast<const ulonglong2 *>(p + r));
        const ulonglong2 b = __ldcs(reinterpret_cast<const ulonglong2 *>(p + r + 2));
        k[0] = a.x; k[1] = a.y; k[2] = b.x; k[3] = b.y;
    }
}

// One iteration of the row loop: U 4-row groups per thread.  CHECK = the groups may straddle [r0, r1).
template <int KW, bool LUT, int NV, int U, bool CHECK>
__device__ __forceinline__ void dagg_step(const DAggParams &P, uint32_t *tab, const typename KRaw<KW>::T *keyp, int64_t g0,
                                          int64_t r0, int64_t r1, uint64_t slot0, int tid) {
    using KT = typename KRaw<KW>::T;
    constexpr int AT = dagg_threads(NV);
    int64_t rr[U];
#pragma unroll
    for (int u = 0; u < U; u++) rr[u] = g0 + (int64_t)(u * AT + tid) * 4;
    KT k[U][4];
    uint32_t x[NV > 0 ? NV : 1][U][4];
#pragma unroll
    for (int u = 0; u < U; u++)
        if (!CHECK || rr[u] < r1) load_keys4<KW>(keyp, rr[u], k[u]);
#pragma unroll
    for (int v = 0; v < NV; v++)
#pragma unroll
        for (int u = 0; u < U; u++)
            if (!CHECK || rr[u] < r1) {
                const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(P.vals[v] + rr[u]));
                x[v][u][0] = q.x; x[v][u][1] = q.y; x[v][u][2] = q.z; x[v][u][3] = q.w;
            }
    uint32_t idx[U][4];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int e = 0; e < 4; e++) {
            idx[u][e] = 0xffffffffu;
            if (!CHECK || (rr[u] + e >= r0 && rr[u] + e < r1)) {
                if (LUT) {
                    long long v;
                    if constexpr (KW == 4) v = (P.key_dtype == HARK_U32 ? (long long)(uint32_t)k[u][e] : (long long)(int32_t)k[u][e]) - P.pk_min;
                    else v = (long long)k[u][e] - P.pk_min;
                    if (v >= 0 && v < P.pk_span) idx[u][e] = __ldg(P.lut + v) - 1u; // 0 (no match) -> 0xffffffff
                } else {
                    idx[u][e] = (uint32_t)(ordkey_of<KW>(k[u][e], P.key_dtype) - P.g_lo - slot0);
                }
            }
        }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int e = 0; e < 4; e++)
            if ((!CHECK && !LUT) || idx[u][e] != 0xffffffffu) atomicAdd(&tab[idx[u][e]], 1u);
#pragma unroll 1
    for (int ai = 0; ai < P.nacc; ai++) {
        const int kind = P.acc[ai].kind, word = P.acc[ai].word, vcol = P.acc[ai].vcol;
#pragma unroll
        for (int v = 0; v < NV; v++)
            if (v == vcol) acc_dispatch<U>(tab, P.K, idx, kind, word, x[v]);
    }
}

template <int KW, bool LUT, int NV>
__global__ void __launch_bounds__(dagg_threads(NV), 1) hk_dagg_kernel(const __grid_constant__ DAggParams P) {
    using KT = typename KRaw<KW>::T;
    constexpr int AT = dagg_threads(NV);
    constexpr int U = NV <= 1 ? 2 : 2; // 4-row groups per thread and iteration (all their loads are in flight together)
    extern __shared__ __align__(16) uint32_t tab[];
    __shared__ unsigned long long s_cpre[258]; // chunks before bucket b
    const int tid = threadIdx.x;
    const uint32_t K = P.K;
    const KT *keyp = reinterpret_cast<const KT *>(P.key);

    // identity-initialise the table
    for (uint32_t i = tid; i < K; i += AT) tab[i] = 0;
    for (int ai = 0; ai < P.nacc; ai++) {
        const DAcc &a = P.acc[ai];
        const int two = (a.kind == A_SUM64S || a.kind == A_SUM64U) ? 2 : 1;
        if (a.kind == A_FSUM || a.kind == A_FPROD) {
            unsigned long long *p = reinterpret_cast<unsigned long long *>(tab + (size_t)a.word * K);
            for (uint32_t i = tid; i < K; i += AT) p[i] = a.kind == A_FPROD ? 0x3ff0000000000000ull : 0ull;
        } else {
            for (int w = 0; w < two; w++)
                for (uint32_t i = tid; i < K; i += AT) tab[(size_t)(a.word + w) * K + i] = acc_identity(a.kind, 0);
        }
    }
    if (tid == 0) {
        unsigned long long run = 0;
        for (int b = 0; b < P.nbins; b++) {
            s_cpre[b] = run;
            const unsigned long long nb = P.offsets[b + 1] - P.offsets[b];
            run += (nb + (unsigned long long)P.chunk - 1) / (unsigned long long)P.chunk;
        }
        s_cpre[P.nbins] = run;
    }
    __syncthreads();
    const long long T = (long long)s_cpre[P.nbins];
    // groupby mode: a contiguous range of chunks per CTA (few bucket changes -> few table merges);
    // lut mode: chunks interleaved over the CTAs, so all CTAs probe the same lookup slice at the same time (L2)
    long long c_begin, c_end, c_step;
    if (LUT) {
        c_begin = blockIdx.x; c_end = T; c_step = gridDim.x;
    } else {
        c_begin = T * (long long)blockIdx.x / (long long)gridDim.x;
        c_end = T * (long long)(blockIdx.x + 1) / (long long)gridDim.x;
        c_step = 1;
    }
    int cur_bucket = -1;
    int b = 0;
    for (long long c = c_begin; c < c_end; c += c_step) {
        while (b + 1 < P.nbins && s_cpre[b + 1] <= (unsigned long long)c) b++; // chunks ascend: a forward walk
        if (!LUT && b != cur_bucket) {
            if (cur_bucket >= 0) {
                __syncthreads();
                for (uint32_t i = tid; i < K; i += AT) flush_slot(tab, K, i, ((uint64_t)cur_bucket << P.shift) + i, P);
                __syncthreads();
            }
            cur_bucket = b;
        }
        if (LUT) cur_bucket = 0;
        const int64_t r0 = (int64_t)P.offsets[b] + (int64_t)(c - (long long)s_cpre[b]) * P.chunk;
        const int64_t r1 = min(r0 + P.chunk, (int64_t)P.offsets[b + 1]);
        const uint64_t slot0 = LUT ? 0ull : ((uint64_t)b << P.shift);
        constexpr int64_t STEP = (int64_t)AT * 4 * U;
        for (int64_t g0 = r0 & ~(int64_t)3; g0 < r1; g0 += STEP) {
            if (g0 >= r0 && g0 + STEP <= r1) dagg_step<KW, LUT, NV, U, false>(P, tab, keyp, g0, r0, r1, slot0, tid);
            else dagg_step<KW, LUT, NV, U, true>(P, tab, keyp, g0, r0, r1, slot0, tid);
        }
    }
    __syncthreads();
    if (cur_bucket >= 0)
        for (uint32_t i = tid; i < K; i += AT) flush_slot(tab, K, i, ((uint64_t)cur_bucket << P.shift) + i, P);
}

// ------------------------------------------------------------------------------------------------
// compaction of the dense accumulators into the output table
// ------------------------------------------------------------------------------------------------
constexpr int CT = 256;
constexpr int CPER = 8;
constexpr int CSPAN = CT * CPER; // dense slots per CTA

__global__ void __launch_bounds__(CT) hk_dense_count_kernel(const unsigned long long *__restrict__ gcnt, uint64_t R,
                                                             uint32_t *__restrict__ block_counts) {
    const uint64_t base = (uint64_t)blockIdx.x * CSPAN + (uint64_t)threadIdx.x * CPER;
    uint32_t c = 0;
#pragma unroll
    for (int e = 0; e < CPER; e++)
        if (base + e < R && gcnt[base + e] != 0) c++;
    c = hk_warp_sum_u32(c);
    __shared__ uint32_t sw[CT / 32];
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < CT / 32; w++) t += sw[w];
        block_counts[blockIdx.x] = t;
    }
}

// exclusive scan of u32 counts -> u64 bases, total in base[count]; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) hk_dense_scan_kernel(const uint32_t *counts, unsigned long long *base, int64_t count) {
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

enum OutKind { O_COUNT = 0, O_COPY32, O_LOW32_OF_U64, O_AVG_S64, O_AVG_U64, O_AVG_F64, O_F32_OF_F64, O_F64, O_F64_OF_S64, O_F64_OF_U64 };
struct OutSpec {
    int kind;
    const void *src; // dense accumulator
    void *dst;       // output column
};
struct CompactParams {
    uint64_t R;
    const unsigned long long *gcnt;
    const unsigned long long *block_base;
    uint64_t g_lo;
    int key_dtype; // output key dtype
    void *out_key;
    int nout;
    OutSpec o[HK_DENSE_MAX_AGGS];
};

__global__ void __launch_bounds__(CT) hk_dense_compact_kernel(const __grid_constant__ CompactParams P) {
    const uint64_t base = (uint64_t)blockIdx.x * CSPAN + (uint64_t)threadIdx.x * CPER;
    unsigned long long cnt[CPER];
    uint32_t c = 0;
#pragma unroll
    for (int e = 0; e < CPER; e++) {
        cnt[e] = base + e < P.R ? P.gcnt[base + e] : 0ull;
        c += cnt[e] != 0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = hk_warp_incl_scan_u32(c);
    __shared__ uint32_t sw[CT / 32];
    if (lane == 31) sw[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += sw[w];
    unsigned long long g = P.block_base[blockIdx.x] + woff + inc - c;
#pragma unroll
    for (int e = 0; e < CPER; e++) {
        if (cnt[e] == 0) continue;
        const uint64_t r = base + e;
        const uint64_t ord = P.g_lo + r;
        switch (P.key_dtype) {
        case HARK_U32: reinterpret_cast<uint32_t *>(P.out_key)[g] = (uint32_t)ord; break;
        case HARK_I32: reinterpret_cast<uint32_t *>(P.out_key)[g] = (uint32_t)ord ^ 0x80000000u; break;
        default: reinterpret_cast<uint64_t *>(P.out_key)[g] = ord ^ 0x8000000000000000ull; break;
        }
        for (int j = 0; j < P.nout; j++) {
            const OutSpec &o = P.o[j];
            switch (o.kind) {
            case O_COUNT: reinterpret_cast<long long *>(o.dst)[g] = (long long)cnt[e]; break;
            case O_COPY32: reinterpret_cast<uint32_t *>(o.dst)[g] = reinterpret_cast<const uint32_t *>(o.src)[r]; break;
            case O_LOW32_OF_U64: reinterpret_cast<uint32_t *>(o.dst)[g] = (uint32_t)reinterpret_cast<const unsigned long long *>(o.src)[r]; break;
            case O_AVG_S64: reinterpret_cast<double *>(o.dst)[g] = (double)reinterpret_cast<const long long *>(o.src)[r] / (double)cnt[e]; break;
            case O_AVG_U64: reinterpret_cast<double *>(o.dst)[g] = (double)reinterpret_cast<const unsigned long long *>(o.src)[r] / (double)cnt[e]; break;
            case O_AVG_F64: reinterpret_cast<double *>(o.dst)[g] = reinterpret_cast<const double *>(o.src)[r] / (double)cnt[e]; break;
            case O_F32_OF_F64: reinterpret_cast<float *>(o.dst)[g] = (float)reinterpret_cast<const double *>(o.src)[r]; break;
            case O_F64_OF_S64: reinterpret_cast<double *>(o.dst)[g] = (double)reinterpret_cast<const long long *>(o.src)[r]; break;
            case O_F64_OF_U64: reinterpret_cast<double *>(o.dst)[g] = (double)reinterpret_cast<const unsigned long long *>(o.src)[r]; break;
            default: reinterpret_cast<double *>(o.dst)[g] = reinterpret_cast<const double *>(o.src)[r]; break;
            }
        }
        g++;
    }
}

template <typename A>
__global__ void __launch_bounds__(256) hk_dense_fill_kernel(A *p, A v, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

unsigned grid_for(hark_ctx *ctx, int64_t n, int per_sm = 8) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * per_sm));
}

template <typename A>
int dfill(hark_ctx *ctx, void *p, A v, uint64_t n) {
    if (n == 0) return HARK_OK;
    hk_dense_fill_kernel<A><<<grid_for(ctx, (int64_t)n), 256, 0, ctx->stream>>>((A *)p, v, n);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch();
    return HARK_OK;
}

struct Bufs { // scratch released on every exit path
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

int floor_log2_u64(uint64_t v) {
    int b = -1;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

} // namespace

int hk_col_minmax(hark_ctx *ctx, const void *col, int32_t dtype, int64_t n, uint64_t *lo, uint64_t *hi) {
    unsigned long long *d = nullptr;
    HK_TRY(ctx->dalloc((void **)&d, 2 * sizeof(unsigned long long)));
    ctx->h_scalars[0] = ~0ull;
    ctx->h_scalars[1] = 0ull;
    cudaError_t e = cudaMemcpyAsync(d, ctx->h_scalars, 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && n > 0) {
        const int kw = hk_dtype_size(dtype);
        const unsigned g = grid_for(ctx, (n + 3) / 4, 8);
        if (kw == 4) hk_dminmax_kernel<4><<<g, 256, 0, ctx->stream>>>(col, n, dtype, d);
        else hk_dminmax_kernel<8><<<g, 256, 0, ctx->stream>>>(col, n, dtype, d);
        e = cudaGetLastError();
        ctx->count_launch();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_scalars, d, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ctx->dfree(d);
    if (e != cudaSuccess) return ctx->fail(HARK_ERR_CUDA, std::string("minmax: ") + cudaGetErrorString(e));
    *lo = ctx->h_scalars[0];
    *hi = ctx->h_scalars[1];
    return HARK_OK;
}

int hk_column_minmax(hark_ctx *ctx, const hark_col &col, int64_t n, int32_t dtype, uint64_t *lo, uint64_t *hi) {
    const bool cacheable = dtype == col.dtype && ctx->opt("stats.cache", 1) != 0;
    if (cacheable && col.mm_valid) {
        *lo = col.mm_lo;
        *hi = col.mm_hi;
        return HARK_OK;
    }
    HK_TRY(hk_col_minmax(ctx, col.ptr, dtype, n, lo, hi));
    if (cacheable) {
        col.mm_lo = *lo;
        col.mm_hi = *hi;
        col.mm_valid = true;
    }
    return HARK_OK;
}

int hk_dense_groupby(hark_ctx *ctx, hark_table **out, const hk_dense_req &rq, bool *handled) {
    *handled = false;
    const int64_t n = rq.n;
    const int kw = hk_dtype_size(rq.key_dtype);
    if (n <= 0 || !hk_dtype_int(rq.key_dtype) || rq.nvals > HK_DENSE_MAX_VALS || rq.c > HK_DENSE_MAX_AGGS) return HARK_OK;
    for (int v = 0; v < rq.nvals; v++)
        if (hk_dtype_size(rq.val_dtypes[v]) != 4) return HARK_OK;
    if (rq.g_hi < rq.g_lo) return HARK_OK;
    const uint64_t Rm1 = rq.g_hi - rq.g_lo;
    if (Rm1 >= (1ull << 40)) return HARK_OK;
    const uint64_t R = Rm1 + 1;

    // ---- accumulators: one per distinct (value column, kind) ----
    struct Acc {
        int vcol, kind, words;
    };
    std::vector<Acc> accs;
    auto acc_index = [&](int vcol, int kind, int words) {
        for (size_t i = 0; i < accs.size(); i++)
            if (accs[i].vcol == vcol && accs[i].kind == kind) return (int)i;
        accs.push_back(Acc{vcol, kind, words});
        return (int)accs.size() - 1;
    };
    // exact 64-bit sums serve both SUM (low word) and AVG; if a column only needs SUM, a 32-bit sum is enough
    std::vector<bool> col_needs_avg((size_t)std::max(rq.nvals, 1), false);
    for (int j = 0; j < rq.c; j++)
        if ((rq.agg_code[j] == HARK_AGG_AVG || rq.agg_code[j] == HARK_AGG_SUMF64) && rq.agg_val[j] >= 0) col_needs_avg[rq.agg_val[j]] = true;
    struct OutPlan {
        int okind, acc;
        int32_t dtype;
    };
    std::vector<OutPlan> outs;
    for (int j = 0; j < rq.c; j++) {
        const int code = rq.agg_code[j], vi = rq.agg_val[j];
        if (code == HARK_AGG_COUNT) {
            outs.push_back({O_COUNT, -1, HARK_I64});
            continue;
        }
        if (vi < 0 || vi >= rq.nvals) return HARK_OK;
        const int32_t vdt = rq.pinned_u32 ? HARK_U32 : rq.val_dtypes[vi];
        const bool is_f = vdt == HARK_F32, is_s = vdt == HARK_I32;
        switch (code) {
        case HARK_AGG_SUM:
            if (is_f) outs.push_back({O_F32_OF_F64, acc_index(vi, A_FSUM, 2), HARK_F32});
            else if (col_needs_avg[vi]) outs.push_back({O_LOW32_OF_U64, acc_index(vi, is_s ? A_SUM64S : A_SUM64U, 2), vdt});
            else outs.push_back({O_COPY32, acc_index(vi, A_SUM32, 1), vdt});
            break;
        case HARK_AGG_AVG:
            if (is_f) outs.push_back({O_AVG_F64, acc_index(vi, A_FSUM, 2), HARK_F64});
            else outs.push_back({is_s ? O_AVG_S64 : O_AVG_U64, acc_index(vi, is_s ? A_SUM64S : A_SUM64U, 2), HARK_F64});
            break;
        case HARK_AGG_SUMF64:
            if (is_f) outs.push_back({O_F64, acc_index(vi, A_FSUM, 2), HARK_F64});
            else outs.push_back({is_s ? O_F64_OF_S64 : O_F64_OF_U64, acc_index(vi, is_s ? A_SUM64S : A_SUM64U, 2), HARK_F64});
            break;
        case HARK_AGG_PROD:
            if (is_f) outs.push_back({O_F32_OF_F64, acc_index(vi, A_FPROD, 2), HARK_F32});
            else outs.push_back({O_COPY32, acc_index(vi, A_PROD32, 1), vdt});
            break;
        case HARK_AGG_MAX:
            outs.push_back({O_COPY32, acc_index(vi, is_f ? A_MAXF : is_s ? A_MAXS : A_MAXU, 1), vdt});
            break;
        default: // MIN, and (pinned entry) code 0 / unknown codes: groupby.fut:41
            outs.push_back({O_COPY32, acc_index(vi, is_f ? A_MINF : is_s ? A_MINS : A_MINU, 1), vdt});
            break;
        }
    }
    if ((int)accs.size() > AMAXACC) return HARK_OK;
    // table words: [count][64-bit accumulators (8-byte aligned pairs)...][32-bit accumulators...]
    int nwords = 1;
    std::vector<int> word_of(accs.size(), 0);
    for (size_t i = 0; i < accs.size(); i++)
        if (accs[i].kind == A_FSUM || accs[i].kind == A_FPROD) {
            nwords += nwords & 1; // u64[K] view needs an even word index (K is a power of two >= 2)
            word_of[i] = nwords;
            nwords += 2;
        }
    for (size_t i = 0; i < accs.size(); i++)
        if (!(accs[i].kind == A_FSUM || accs[i].kind == A_FPROD)) {
            word_of[i] = nwords;
            nwords += accs[i].words;
        }
    const int64_t smem_budget = ctx->opt("dense.smem_bytes", 200 * 1024);
    int shift = floor_log2_u64((uint64_t)smem_budget / (4ull * (uint64_t)nwords));
    const int64_t force_shift = ctx->opt("dense.log2_slots", 0); // tests: small tables exercise the partition pass
    if (force_shift > 0) shift = (int)std::min<int64_t>(shift, force_shift);
    if (shift < 8) return HARK_OK;
    const uint64_t K = 1ull << shift;
    const bool lut_mode = rq.lut != nullptr;
    const uint64_t nbins64 = (R + K - 1) >> shift;
    if (lut_mode ? (R > K) : (nbins64 > 256)) return HARK_OK;
    if (!lut_mode && R > 16ull * (uint64_t)n + (1ull << 16)) return HARK_OK; // sparse keys: dense slots would dominate
    const int nbins = lut_mode ? 1 : (int)nbins64;
    if (nbins > 1 && rq.nvals > PMAXV) return HARK_OK;
    *handled = true;

    Bufs scratch(ctx);
    ctx->kernel_begin(); // kernel_ms = partition pass (if any) + aggregation pass
    // ---- bucket offsets (+ partition pass when the key range needs more than one table) ----
    const void *key = rq.key;
    const void *vals[HK_DENSE_MAX_VALS] = {rq.vals[0], rq.vals[1], rq.vals[2], rq.vals[3]};
    unsigned long long *d_offsets = nullptr;
    int agg_nbins = nbins;
    // lut mode: slice the lookup so that the slice being probed stays L2-resident
    int lut_bins = 1, lut_shift = 0;
    if (lut_mode) {
        const int64_t slice_bytes = ctx->opt("join.lut_slice_bytes", 16ll << 20);
        const uint64_t lut_bytes = (uint64_t)rq.pk_span * 4ull;
        if ((int64_t)lut_bytes > slice_bytes * 3 / 2 && rq.nvals <= PMAXV) {
            lut_shift = floor_log2_u64((uint64_t)slice_bytes / 4);
            while ((((uint64_t)rq.pk_span - 1) >> lut_shift) + 1 > 256) lut_shift++;
            lut_bins = (int)((((uint64_t)rq.pk_span - 1) >> lut_shift) + 1);
        }
    }
    if (nbins > 1 || lut_bins > 1) {
        hk_part_spec ps;
        if (lut_mode) {
            // order key of pk_min in the fact key's dtype (signed dtypes: sign-bit flip)
            ps.dtype = rq.key_dtype;
            ps.base = kw == 4 ? (uint64_t)(rq.key_dtype == HARK_U32 ? (uint32_t)rq.pk_min : ((uint32_t)(int32_t)rq.pk_min ^ 0x80000000u))
                              : ((uint64_t)rq.pk_min ^ 0x8000000000000000ull);
            ps.span = (uint64_t)rq.pk_span;
            ps.shift = lut_shift;
            ps.nbins = lut_bins;
            agg_nbins = lut_bins;
        } else {
            ps.dtype = rq.key_dtype;
            ps.base = rq.g_lo;
            ps.span = R;
            ps.shift = shift;
            ps.nbins = nbins;
        }
        void *ko = nullptr, *vo[PMAXV] = {nullptr, nullptr, nullptr};
        HK_TRY(hk_partition_pass(ctx, n, key, kw, ps, rq.nvals, vals, &ko, vo, &d_offsets));
        scratch.adopt(ko);
        scratch.adopt(d_offsets);
        key = ko;
        for (int v = 0; v < rq.nvals; v++) {
            scratch.adopt(vo[v]);
            vals[v] = vo[v];
        }
    } else {
        HK_TRY(scratch.alloc((void **)&d_offsets, 2 * sizeof(unsigned long long)));
        ctx->h_scalars[8] = 0;
        ctx->h_scalars[9] = (uint64_t)n;
        HK_CUDA(ctx, cudaMemcpyAsync(d_offsets, ctx->h_scalars + 8, 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    }

    // ---- dense global accumulators ----
    unsigned long long *gcnt = nullptr;
    HK_TRY(scratch.alloc((void **)&gcnt, R * sizeof(unsigned long long)));
    HK_CUDA(ctx, cudaMemsetAsync(gcnt, 0, R * sizeof(unsigned long long), ctx->stream));
    DAggParams P;
    memset(&P, 0, sizeof P);
    P.key = key;
    P.key_dtype = rq.key_dtype;
    P.n = n;
    P.g_lo = rq.g_lo;
    P.shift = shift;
    P.K = (uint32_t)K;
    P.nwords = nwords;
    P.nbins = agg_nbins;
    P.offsets = d_offsets;
    P.lut = rq.lut;
    P.pk_min = rq.pk_min;
    P.pk_span = rq.pk_span;
    P.nvals = rq.nvals;
    for (int v = 0; v < rq.nvals; v++) P.vals[v] = (const uint32_t *)vals[v];
    P.nacc = (int)accs.size();
    P.gcnt = gcnt;
    std::vector<void *> gacc(accs.size(), nullptr);
    for (size_t i = 0; i < accs.size(); i++) {
        const int kind = accs[i].kind;
        const bool wide = kind == A_SUM64S || kind == A_SUM64U || kind == A_FSUM || kind == A_FPROD;
        HK_TRY(scratch.alloc(&gacc[i], R * (wide ? 8 : 4)));
        switch (kind) {
        case A_FPROD: HK_TRY(dfill<double>(ctx, gacc[i], 1.0, R)); break;
        case A_MINU: HK_TRY(dfill<uint32_t>(ctx, gacc[i], 0xffffffffu, R)); break;
        case A_MINS: HK_TRY(dfill<uint32_t>(ctx, gacc[i], 0x7fffffffu, R)); break;
        case A_MAXS: HK_TRY(dfill<uint32_t>(ctx, gacc[i], 0x80000000u, R)); break;
        case A_MINF: HK_TRY(dfill<uint32_t>(ctx, gacc[i], 0x7f800000u, R)); break;
        case A_MAXF: HK_TRY(dfill<uint32_t>(ctx, gacc[i], 0xff800000u, R)); break;
        case A_PROD32: HK_TRY(dfill<uint32_t>(ctx, gacc[i], 1u, R)); break;
        default: HK_CUDA(ctx, cudaMemsetAsync(gacc[i], 0, R * (wide ? 8 : 4), ctx->stream)); break;
        }
        P.acc[i] = DAcc{accs[i].vcol, kind, word_of[i], gacc[i]};
    }
    const unsigned grid = (unsigned)ctx->num_sms;
    {
        int64_t chunk = (n + (int64_t)grid * 8 - 1) / ((int64_t)grid * 8);
        chunk = std::max<int64_t>(chunk, 16384);
        P.chunk = (chunk + 3) & ~(int64_t)3;
    }
    const size_t smem = (size_t)K * 4 * nwords;
    {
        cudaError_t e;
        void (*kern)(const DAggParams) = nullptr;
#define HK_DAGG_PICK(KWv, LUTv)                                                        \
    switch (rq.nvals) {                                                                \
    case 0: kern = hk_dagg_kernel<KWv, LUTv, 0>; break;                                \
    case 1: kern = hk_dagg_kernel<KWv, LUTv, 1>; break;                                \
    case 2: kern = hk_dagg_kernel<KWv, LUTv, 2>; break;                                \
    case 3: kern = hk_dagg_kernel<KWv, LUTv, 3>; break;                                \
    default: kern = hk_dagg_kernel<KWv, LUTv, 4>; break;                               \
    }
        if (lut_mode) {
            if (kw == 4) { HK_DAGG_PICK(4, true) } else { HK_DAGG_PICK(8, true) }
        } else {
            if (kw == 4) { HK_DAGG_PICK(4, false) } else { HK_DAGG_PICK(8, false) }
        }
#undef HK_DAGG_PICK
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) kern<<<grid, dagg_threads(rq.nvals), smem, ctx->stream>>>(P);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return ctx->fail(HARK_ERR_CUDA, std::string("dense aggregate: ") + cudaGetErrorString(e));
        ctx->count_launch();
    }
    ctx->kernel_end();

    // ---- compaction: dense slots -> output rows (ascending slot = ascending key) ----
    const int64_t nblk = (int64_t)((R + CSPAN - 1) / CSPAN);
    uint32_t *bc = nullptr;
    unsigned long long *bb = nullptr;
    HK_TRY(scratch.alloc((void **)&bc, sizeof(uint32_t) * (size_t)nblk));
    HK_TRY(scratch.alloc((void **)&bb, sizeof(unsigned long long) * (size_t)(nblk + 1)));
    hk_dense_count_kernel<<<(unsigned)nblk, CT, 0, ctx->stream>>>(gcnt, R, bc);
    HK_CHECK_LAUNCH(ctx);
    hk_dense_scan_kernel<<<1, 1024, 0, ctx->stream>>>(bc, bb, nblk);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch(2);
    HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, bb + nblk, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t G = (int64_t)ctx->h_scalars[0];

    std::vector<int32_t> odt(1 + (size_t)rq.c);
    odt[0] = rq.out_key_dtype;
    for (int j = 0; j < rq.c; j++) odt[1 + j] = rq.pinned_u32 ? HARK_U32 : outs[j].dtype;
    hark_table *t = nullptr;
    HK_TRY(hk_table_alloc(ctx, &t, G, G, odt.data(), 1 + rq.c));
    if (G > 0) {
        CompactParams C;
        memset(&C, 0, sizeof C);
        C.R = R;
        C.gcnt = gcnt;
        C.block_base = bb;
        C.g_lo = rq.g_lo;
        C.key_dtype = rq.out_key_dtype;
        C.out_key = t->cols[0].ptr;
        C.nout = rq.c;
        for (int j = 0; j < rq.c; j++) C.o[j] = OutSpec{outs[j].okind, outs[j].acc >= 0 ? gacc[outs[j].acc] : nullptr, t->cols[1 + j].ptr};
        hk_dense_compact_kernel<<<(unsigned)nblk, CT, 0, ctx->stream>>>(C);
        cudaError_t e = cudaGetLastError();
        ctx->count_launch();
        if (e != cudaSuccess) {
            hark_table_free(ctx, t);
            return ctx->fail(HARK_ERR_CUDA, std::string("dense compact: ") + cudaGetErrorString(e));
        }
    }
    *out = t;
    return HARK_OK;
}
