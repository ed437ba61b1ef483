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

// Zone map of an f32 column for exact fixed-point sums: out[0] = max exponent of the highest set bit, out[1] = min
// exponent of the lowest set bit (both over non-zero values, biased by +1024 so that they fit unsigned atomics),
// out[2] = number of non-finite values, out[3] = number of non-zero values.
__global__ void __launch_bounds__(256) hk_f32_fxstats_kernel(const uint32_t *__restrict__ col, int64_t n, unsigned long long *out) {
    int hi = -100000, lo = 100000;
    unsigned long long bad = 0, any = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t b = __ldcs(col + i);
        const uint32_t ex = (b >> 23) & 0xffu, fr = b & 0x7fffffu;
        if (ex == 0xffu) {
            bad++;
            continue;
        }
        if (ex == 0u && fr == 0u) continue;
        const uint32_t mant = ex ? (fr | 0x800000u) : fr;
        const int e0 = (ex ? (int)ex : 1) - 127 - 23; // exponent of mantissa bit 0
        hi = max(hi, e0 + 31 - __clz((int)mant));
        lo = min(lo, e0 + __ffs((int)mant) - 1);
        any++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        hi = max(hi, __shfl_xor_sync(HK_FULL_MASK, hi, o));
        lo = min(lo, __shfl_xor_sync(HK_FULL_MASK, lo, o));
        bad += __shfl_xor_sync(HK_FULL_MASK, bad, o);
        any += __shfl_xor_sync(HK_FULL_MASK, any, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (any) {
            atomicMax(out, (unsigned long long)(hi + 1024));
            atomicMin(out + 1, (unsigned long long)(lo + 1024));
            atomicAdd(out + 3, any);
        }
        if (bad) atomicAdd(out + 2, bad);
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
    A_PROD32,     // u32 wrap-around product
    A_FXSUM32,    // f32 -> exact fixed point, |v| / 2^lo < 2^31: one 32-bit addend per row (table and global as A_SUM64S)
    A_FXSUM64     // f32 -> exact fixed point with 64-bit addends (two table atomics per row)
};

struct DAcc {
    int vcol;
    int kind;
    int word;     // first table word (units of K u32)
    void *gacc;   // dense global accumulator [R]
    float fscale; // A_FXSUM*: 2^-lo, the value's fixed-point representation is v * fscale (an integer, exactly)
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
    // K8t input (hk_dagg_tiles_kernel): the table as tile blocks sorted by bin + a directory (partition.cu)
    const uint32_t *t_rows;
    const uint32_t *t_dir;
    long long t_num_tiles;
    const unsigned long long *t_cum; // MODE 0: exclusive row prefix over (bin, block of TBLK tiles), bin-major, [nbins * t_nblk + 1]
    long long t_nblk;
    // hash mode (join + GROUP BY over sparse keys, join.cu): open addressing, slot + 1 in the entry, 0 = empty
    const void *htab;
    unsigned long long hmask;
    unsigned long long *ticket; // lookup / hash modes: work units are dealt from this counter (zeroed by the host)
    unsigned long long *bin_ticket; // GROUP BY mode: one counter per bin (zeroed by the host); null = static row split
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
__device__ __forceinline__ void acc_row(uint32_t *tab, uint32_t K, uint32_t idx, int word, uint32_t x, uint32_t xhi = 0) {
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
    } else if constexpr (KIND == A_FXSUM64) { // x = low word here; the high word travels in `xhi`
        const uint32_t old = atomicAdd(w0, x);
        atomicAdd(w0 + K, xhi + (uint32_t)((uint32_t)(old + x) < old));
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

// f32 bits -> the exact fixed-point integer v * fscale
__device__ __forceinline__ long long fx_of(uint32_t bits, float fscale) { return __float2ll_rn(__uint_as_float(bits) * fscale); }

template <int U>
__device__ __forceinline__ void acc_dispatch(uint32_t *tab, uint32_t K, const uint32_t (&idx)[U][4], int kind, int word,
                                             const uint32_t (&x)[U][4], float fscale) {
    switch (kind) {
    case A_FXSUM32:
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (idx[u][e] != 0xffffffffu) acc_row<A_SUM64S>(tab, K, idx[u][e], word, (uint32_t)(int32_t)fx_of(x[u][e], fscale));
        break;
    case A_FXSUM64:
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (idx[u][e] != 0xffffffffu) {
                    const unsigned long long v = (unsigned long long)fx_of(x[u][e], fscale);
                    acc_row<A_FXSUM64>(tab, K, idx[u][e], word, (uint32_t)v, (uint32_t)(v >> 32));
                }
        break;
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
        case A_SUM64U:
        case A_FXSUM32:
        case A_FXSUM64: {
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

template <int KW>
__device__ __forceinline__ uint32_t hash_probe(const void *htab, uint64_t hmask, typename KRaw<KW>::T key);

// One iteration of the row loop: U 4-row groups per thread.  CHECK = the groups may straddle [r0, r1).
// MODE: 0 slot = key - first key of the bucket, 1 slot from the direct-address lookup, 2 slot from the hash table.
template <int KW, int MODE, int NV, int U, bool CHECK>
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
                if constexpr (MODE == 1) {
                    long long v;
                    if constexpr (KW == 4) v = (P.key_dtype == HARK_U32 ? (long long)(uint32_t)k[u][e] : (long long)(int32_t)k[u][e]) - P.pk_min;
                    else v = (long long)k[u][e] - P.pk_min;
                    if (v >= 0 && v < P.pk_span) idx[u][e] = __ldg(P.lut + v) - 1u; // 0 (no match) -> 0xffffffff
                } else if constexpr (MODE == 2) {
                    idx[u][e] = hash_probe<KW>(P.htab, P.hmask, k[u][e]);
                } else {
                    idx[u][e] = (uint32_t)(ordkey_of<KW>(k[u][e], P.key_dtype) - P.g_lo - slot0);
                }
            }
        }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int e = 0; e < 4; e++)
            if ((!CHECK && MODE == 0) || idx[u][e] != 0xffffffffu) atomicAdd(&tab[idx[u][e]], 1u);
#pragma unroll 1
    for (int ai = 0; ai < P.nacc; ai++) {
        const int kind = P.acc[ai].kind, word = P.acc[ai].word, vcol = P.acc[ai].vcol;
#pragma unroll
        for (int v = 0; v < NV; v++)
            if (v == vcol) acc_dispatch<U>(tab, P.K, idx, kind, word, x[v], P.acc[ai].fscale);
    }
}

template <int KW, int MODE, int NV>
__global__ void __launch_bounds__(dagg_threads(NV), 1) hk_dagg_kernel(const __grid_constant__ DAggParams P) {
    constexpr bool LUT = MODE != 0;
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
        const int two = (a.kind == A_SUM64S || a.kind == A_SUM64U || a.kind == A_FXSUM32 || a.kind == A_FXSUM64) ? 2 : 1;
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
            if (g0 >= r0 && g0 + STEP <= r1) dagg_step<KW, MODE, NV, U, false>(P, tab, keyp, g0, r0, r1, slot0, tid);
            else dagg_step<KW, MODE, NV, U, true>(P, tab, keyp, g0, r0, r1, slot0, tid);
        }
    }
    __syncthreads();
    if (cur_bucket >= 0)
        for (uint32_t i = tid; i < K; i += AT) flush_slot(tab, K, i, ((uint64_t)cur_bucket << P.shift) + i, P);
}

// ------------------------------------------------------------------------------------------------
// K2 over K8t's tile blocks.  Work unit = the run of one bin inside one tile (a "segment", on average
// 4096 / nbins rows, contiguous, array-of-structs); a warp takes TG segments of the same bin at a time and keeps the
// directory words of all of them, then 2 row loads per segment and lane, in flight together.
//   MODE 0 (GROUP BY)  units are bin-major; every CTA owns a contiguous range of them, so it changes bin (and merges
//                      its table into the global accumulators) at most a few times; slot = key - bin's first key;
//   MODE 1 (lookup)    slot + 1 = lut[fk - pk_min]; units are dealt round-robin in bin-major order so that all CTAs
//                      probe the same slice of the lookup at the same time (the slice stays L2-resident);
//   MODE 2 (hash)      slot + 1 = the entry of an open-addressing table keyed by hk_hash_key (join.cu), same dealing.
// ------------------------------------------------------------------------------------------------
constexpr int TG = 8;   // segments (tiles of one bin) per warp and iteration

template <int KW>
__device__ __forceinline__ uint32_t hash_probe(const void *htab, uint64_t hmask, typename KRaw<KW>::T key) {
    uint64_t h = hk_hash_key<KW>(key) & hmask;
    if constexpr (KW == 4) {
        const uint2 *t = reinterpret_cast<const uint2 *>(htab); // {key, slot + 1}
        while (true) {
            const uint2 e = __ldg(t + h);
            if (e.y == 0u) return 0xffffffffu;
            if (e.x == key) return e.y - 1u;
            h = (h + 1) & hmask;
        }
    } else {
        const ulonglong2 *t = reinterpret_cast<const ulonglong2 *>(htab); // {key, slot + 1 | row << 32}
        while (true) {
            const ulonglong2 e = __ldg(t + h);
            if ((uint32_t)e.y == 0u) return 0xffffffffu;
            if (e.x == key) return (uint32_t)e.y - 1u;
            h = (h + 1) & hmask;
        }
    }
}

template <int KIND, int N>
__device__ __forceinline__ void acc_rows1(uint32_t *tab, uint32_t K, const uint32_t (&idx)[N], int word, const uint32_t (&x)[N]) {
#pragma unroll
    for (int u = 0; u < N; u++)
        if (idx[u] != 0xffffffffu) acc_row<KIND>(tab, K, idx[u], word, x[u]);
}

template <int N>
__device__ __forceinline__ void acc_dispatch1(uint32_t *tab, uint32_t K, const uint32_t (&idx)[N], int kind, int word,
                                              const uint32_t (&x)[N], float fscale) {
    switch (kind) {
    case A_FXSUM32:
#pragma unroll
        for (int u = 0; u < N; u++)
            if (idx[u] != 0xffffffffu) acc_row<A_SUM64S>(tab, K, idx[u], word, (uint32_t)(int32_t)fx_of(x[u], fscale));
        break;
    case A_FXSUM64:
#pragma unroll
        for (int u = 0; u < N; u++)
            if (idx[u] != 0xffffffffu) {
                const unsigned long long v = (unsigned long long)fx_of(x[u], fscale);
                acc_row<A_FXSUM64>(tab, K, idx[u], word, (uint32_t)v, (uint32_t)(v >> 32));
            }
        break;
    case A_SUM32: acc_rows1<A_SUM32, N>(tab, K, idx, word, x); break;
    case A_SUM64S: acc_rows1<A_SUM64S, N>(tab, K, idx, word, x); break;
    case A_SUM64U: acc_rows1<A_SUM64U, N>(tab, K, idx, word, x); break;
    case A_FSUM: acc_rows1<A_FSUM, N>(tab, K, idx, word, x); break;
    case A_FPROD: acc_rows1<A_FPROD, N>(tab, K, idx, word, x); break;
    case A_MINU: acc_rows1<A_MINU, N>(tab, K, idx, word, x); break;
    case A_MAXU: acc_rows1<A_MAXU, N>(tab, K, idx, word, x); break;
    case A_MINS: acc_rows1<A_MINS, N>(tab, K, idx, word, x); break;
    case A_MAXS: acc_rows1<A_MAXS, N>(tab, K, idx, word, x); break;
    case A_MINF: acc_rows1<A_MINF, N>(tab, K, idx, word, x); break;
    case A_MAXF: acc_rows1<A_MAXF, N>(tab, K, idx, word, x); break;
    default: acc_rows1<A_PROD32, N>(tab, K, idx, word, x); break;
    }
}

// SPEC: the accumulator programme, decided on the host so that the row loop carries no run-time switch for the
// common shapes: 0 generic (any list of accumulators), 1 COUNT only, 2 COUNT + exact 64-bit sum of i32 values of value
// column 0 (SUM + COUNT + AVG of one column: config 3), 3 COUNT + 32-bit wrap-around sum (config 5), 4 COUNT + exact
// fixed-point sum of f32 values (config 3, f32 variant).  Specialised programmes keep their accumulator at table word 1.
enum { SPEC_GENERIC = 0, SPEC_COUNT = 1, SPEC_SUM64S = 2, SPEC_SUM32 = 3, SPEC_FX32 = 4 };

// TG segments of bin `b` (tiles t0 .. t0+TG-1, those < t_end) by one warp.  The segments differ in length by a few
// rows, so the lanes walk them through a common WINDOW: virtual row v = g * W + off (W = longest of the TG runs, at
// least 32) belongs to segment g, offset off; off >= length is a hole.  Lanes 0..TG-1 hold the directory words, a row
// fetches its segment's (start, length) with one indexed shuffle.  Lane utilisation = mean / max run length (~85 % for
// 64 bins) where a per-segment loop in steps of 32 rows reaches 41 % (ncu, round 2).
template <int KW, int MODE, int NV, int SPEC>
__device__ __forceinline__ void dagg_segments(const DAggParams &P, uint32_t *tab, int b, long long t0, long long t_end, int lane) {
    using KT = typename KRaw<KW>::T;
    constexpr int RW = KW / 4 + NV;
    constexpr int U = RW <= 2 ? 8 : 4; // rows per lane in flight
    uint32_t w = 0;
    if (lane < TG && t0 + lane < t_end) w = __ldg(P.t_dir + (size_t)(t0 + lane) * P.nbins + b);
    const uint32_t seg_s = w & 0xffffu, seg_len = (w >> 16) - seg_s;
    const uint32_t packed = seg_s | (seg_len << 16);
    uint32_t W = seg_len;
    W = max(W, __shfl_xor_sync(HK_FULL_MASK, W, 1));
    W = max(W, __shfl_xor_sync(HK_FULL_MASK, W, 2));
    W = max(W, __shfl_xor_sync(HK_FULL_MASK, W, 4));
    W = __shfl_sync(HK_FULL_MASK, W, 0);
    if (W == 0) return;
    W = max(W, 32u);
    const uint32_t total = W * TG;
    const uint32_t *base = P.t_rows + (size_t)t0 * HK_TPART_TILE * RW;
    const uint32_t K = P.K;
    const uint32_t slot0 = MODE == 0 ? (uint32_t)(P.g_lo + ((uint64_t)b << P.shift)) : 0u; // low word is enough: slots < K
    const uint64_t slot0_64 = MODE == 0 ? (P.g_lo + ((uint64_t)b << P.shift)) : 0ull;
    uint32_t g = 0, off = (uint32_t)lane; // lane's current virtual row (W >= 32 > lane)
    for (uint32_t v0 = 0; v0 < total; v0 += 32 * U) {
        KT k[U];
        uint32_t x[NV > 0 ? NV : 1][U];
        uint32_t okm = 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t pk = __shfl_sync(HK_FULL_MASK, packed, (int)g); // lanes >= TG hold 0: past the last segment
            k[u] = 0;
            if (off < (pk >> 16)) {
                okm |= 1u << u;
                const uint32_t *row = base + (size_t)(g * HK_TPART_TILE + (pk & 0xffffu) + off) * RW;
                if constexpr (RW == 1) {
                    k[u] = __ldcs(row);
                } else if constexpr (RW == 2) {
                    const uint2 q = __ldcs(reinterpret_cast<const uint2 *>(row));
                    if constexpr (KW == 4) { k[u] = q.x; x[0][u] = q.y; }
                    else k[u] = (uint64_t)q.x | ((uint64_t)q.y << 32);
                } else if constexpr (RW == 4) {
                    const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(row));
                    if constexpr (KW == 4) { k[u] = q.x; x[0][u] = q.y; x[1][u] = q.z; x[2][u] = q.w; }
                    else { k[u] = (uint64_t)q.x | ((uint64_t)q.y << 32); x[0][u] = q.z; x[1][u] = q.w; }
                } else {
                    uint32_t ww[RW];
#pragma unroll
                    for (int i = 0; i < RW; i++) ww[i] = __ldcs(row + i);
                    if constexpr (KW == 4) k[u] = ww[0];
                    else k[u] = (uint64_t)ww[0] | ((uint64_t)ww[1] << 32);
#pragma unroll
                    for (int v = 0; v < NV; v++) x[v][u] = ww[KW / 4 + v];
                }
            }
            off += 32;
            if (off >= W) {
                off -= W;
                g++;
            }
        }
        uint32_t idx[U];
        if constexpr (MODE == 2) {
            // hash probes: the FIRST probe of every row is issued before any is looked at (U independent loads in
            // flight, like the lookup mode); only rows whose first entry holds another key walk on
            if constexpr (KW == 4) {
                // 8-byte entries: a probe reads the 16-byte PAIR that holds its slot (the load costs the same L2 sector as
                // one entry would) and settles up to two probe positions per load; the first pair of every row is requested
                // before any is examined.  Fewer trips through the walk matters twice: a warp repeats the loop until its
                // slowest lane is done (linear probing at load 0.37: ~6 single-entry trips per warp, ~3 pair trips).
                const uint4 *ht = reinterpret_cast<const uint4 *>(P.htab);
                const uint64_t pmask = P.hmask >> 1;
                uint64_t h[U];
                uint4 e0[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    h[u] = hk_hash_key<4>(k[u]) & P.hmask;
                    e0[u] = make_uint4(0u, 0u, 0u, 0u);
                    if ((okm >> u) & 1u) e0[u] = __ldg(ht + (h[u] >> 1));
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    idx[u] = 0xffffffffu;
                    if ((okm >> u) & 1u) {
                        uint4 e = e0[u];
                        uint64_t pp = h[u] >> 1;
                        bool second_only = (h[u] & 1ull) != 0; // the slot is the pair's second entry: the first is not on the probe path
                        while (true) {
                            if (!second_only) {
                                if (e.y == 0u) break;
                                if (e.x == k[u]) {
                                    idx[u] = e.y - 1u;
                                    break;
                                }
                            }
                            if (e.w == 0u) break;
                            if (e.z == k[u]) {
                                idx[u] = e.w - 1u;
                                break;
                            }
                            second_only = false;
                            pp = (pp + 1) & pmask;
                            e = __ldg(ht + pp);
                        }
                    }
                }
            } else {
                const ulonglong2 *ht = reinterpret_cast<const ulonglong2 *>(P.htab);
                uint64_t h[U];
                ulonglong2 e0[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    h[u] = hk_hash_key<8>(k[u]) & P.hmask;
                    e0[u] = make_ulonglong2(0ull, 0ull);
                    if ((okm >> u) & 1u) e0[u] = __ldg(ht + h[u]);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    idx[u] = 0xffffffffu;
                    if ((okm >> u) & 1u) {
                        ulonglong2 e = e0[u];
                        uint64_t hh = h[u];
                        while ((uint32_t)e.y != 0u && e.x != k[u]) {
                            hh = (hh + 1) & P.hmask;
                            e = __ldg(ht + hh);
                        }
                        if ((uint32_t)e.y != 0u) idx[u] = (uint32_t)e.y - 1u;
                    }
                }
            }
        } else {
#pragma unroll
        for (int u = 0; u < U; u++) {
            idx[u] = 0xffffffffu;
            if ((okm >> u) & 1u) {
                if constexpr (MODE == 1) {
                    long long v;
                    if constexpr (KW == 4) v = (P.key_dtype == HARK_U32 ? (long long)(uint32_t)k[u] : (long long)(int32_t)k[u]) - P.pk_min;
                    else v = (long long)k[u] - P.pk_min;
                    if (v >= 0 && v < P.pk_span) idx[u] = __ldg(P.lut + v) - 1u;
                } else if constexpr (KW == 4) {
                    idx[u] = hk_ordkey32(k[u], P.key_dtype) - slot0;
                } else {
                    idx[u] = (uint32_t)(hk_ordkey64(k[u], P.key_dtype) - slot0_64);
                }
            }
        }
        }
        if constexpr (SPEC == SPEC_GENERIC) {
#pragma unroll
            for (int u = 0; u < U; u++)
                if (idx[u] != 0xffffffffu) atomicAdd(&tab[idx[u]], 1u);
#pragma unroll 1
            for (int ai = 0; ai < P.nacc; ai++) {
                const int kind = P.acc[ai].kind, word = P.acc[ai].word, vcol = P.acc[ai].vcol;
#pragma unroll
                for (int v = 0; v < NV; v++)
                    if (v == vcol) acc_dispatch1<U>(tab, K, idx, kind, word, x[v], P.acc[ai].fscale);
            }
        } else {
            [[maybe_unused]] const float fscale = P.acc[0].fscale;
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (idx[u] == 0xffffffffu) continue;
                atomicAdd(&tab[idx[u]], 1u);
                if constexpr (SPEC == SPEC_SUM32) {
                    atomicAdd(&tab[K + idx[u]], x[0][u]);
                } else if constexpr (SPEC == SPEC_SUM64S || SPEC == SPEC_FX32) {
                    const uint32_t xv = SPEC == SPEC_FX32 ? (uint32_t)(int32_t)fx_of(x[0][u], fscale) : x[0][u];
                    const uint32_t old = atomicAdd(&tab[K + idx[u]], xv);
                    const int delta = (int)((uint32_t)(old + xv) < old) - (int)((int32_t)xv < 0);
                    if (delta != 0) atomicAdd(reinterpret_cast<int *>(&tab[2 * K + idx[u]]), delta);
                }
            }
        }
    }
}

// Rows of every (bin, block of TBLK tiles): what the GROUP BY mode balances its CTAs by (bins of skewed keys differ in
// size by orders of magnitude).  Consecutive threads take consecutive bins of one block, so the directory reads coalesce.
constexpr int TBLK = 64;
static_assert(TBLK % TG == 0, "a block is a whole number of warp units");

__global__ void __launch_bounds__(256) hk_tile_block_rows_kernel(const uint32_t *__restrict__ dir, long long num_tiles, int nbins,
                                                                  long long nblk, uint32_t *__restrict__ blk_rows) {
    const long long total = nblk * nbins;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long blk = i / nbins;
        const int b = (int)(i - blk * nbins);
        const long long t1 = min(num_tiles, (blk + 1) * TBLK);
        uint32_t sum = 0;
        for (long long t = blk * TBLK; t < t1; t++) {
            const uint32_t w = __ldg(dir + (size_t)t * nbins + b);
            sum += (w >> 16) - (w & 0xffffu);
        }
        blk_rows[(size_t)b * nblk + blk] = sum;
    }
}

template <int KW, int MODE, int NV, int SPEC>
__global__ void __launch_bounds__(dagg_threads(NV), 1) hk_dagg_tiles_kernel(const __grid_constant__ DAggParams P) {
    constexpr int AT = dagg_threads(NV);
    constexpr int NW = AT / 32;
    extern __shared__ __align__(16) uint32_t tab[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t K = P.K;

    for (uint32_t i = tid; i < K; i += AT) tab[i] = 0;
    for (int ai = 0; ai < P.nacc; ai++) {
        const DAcc &a = P.acc[ai];
        const int two = (a.kind == A_SUM64S || a.kind == A_SUM64U || a.kind == A_FXSUM32 || a.kind == A_FXSUM64) ? 2 : 1;
        if (a.kind == A_FSUM || a.kind == A_FPROD) {
            unsigned long long *p = reinterpret_cast<unsigned long long *>(tab + (size_t)a.word * K);
            for (uint32_t i = tid; i < K; i += AT) p[i] = a.kind == A_FPROD ? 0x3ff0000000000000ull : 0ull;
        } else {
            for (int w = 0; w < two; w++)
                for (uint32_t i = tid; i < K; i += AT) tab[(size_t)(a.word + w) * K + i] = acc_identity(a.kind, 0);
        }
    }
    __syncthreads();
    const long long NT = P.t_num_tiles;
    const long long groups_per_bin = (NT + TG - 1) / TG;            // unit = TG tiles of one bin
    const long long total = groups_per_bin * (long long)P.nbins;
    if constexpr (MODE == 0) {
        // this CTA's share of the ROWS, in (bin, block) entries of the bin-major prefix P.t_cum
        __shared__ long long s_e[2];
        const long long NB = P.t_nblk, nent = NB * (long long)P.nbins;
        if (P.t_cum == nullptr) { // dealt units without the row prefix: homes spread evenly over the bins
            if (tid < 2) s_e[tid] = (long long)(blockIdx.x + tid) * nent / gridDim.x;
        } else if (tid < 2) {
            const unsigned long long rows_total = P.t_cum[nent];
            const unsigned long long target = rows_total / gridDim.x * (blockIdx.x + tid) +
                                              rows_total % gridDim.x * (blockIdx.x + tid) / gridDim.x;
            long long lo = 0, hi = nent; // first entry whose exclusive prefix is >= target
            if (blockIdx.x + tid == gridDim.x) lo = nent;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (P.t_cum[mid] < target) lo = mid + 1; else hi = mid;
            }
            s_e[tid] = lo;
        }
        __syncthreads();
        const long long e0 = s_e[0], e1 = s_e[1];
        if (P.bin_ticket) {
            // Dynamic dealing (default).  The static split below gives every CTA the same number of ROWS, but rows do not
            // cost the same: with Zipf keys the bin that holds the hot groups serialises on shared-memory atomics while
            // the cold bins' short segments waste lanes, and the chip waited for a few CTAs (ncu: SMs busy 27 % of the
            // kernel, 11 ms where the uniform table takes 2.8, profiles/r02_ad_zipf_dagg_ncu.txt).  Every bin has a
            // ticket counter dealing units of TG tiles; a CTA starts on the bin its row share starts in, drains it with
            // all its warps, merges its table, and moves on to the next bin that still has units — so the CTAs of cheap
            // bins end up helping with the expensive one.  A warp asks for its next ticket before it works on the
            // current one (the atomic's latency is hidden).
            const int nb = P.nbins;
            const int home = (int)min((long long)nb - 1, e0 / NB);
            for (int step = 0; step < nb; step++) {
                const int b = home + step < nb ? home + step : home + step - nb;
                unsigned long long *tk = P.bin_ticket + b;
                bool did = false;
                long long un = 0;
                if (lane == 0) un = (long long)atomicAdd(tk, 1ull);
                while (true) {
                    const long long t = __shfl_sync(HK_FULL_MASK, un, 0) * TG;
                    if (t >= NT) break;
                    if (lane == 0) un = (long long)atomicAdd(tk, 1ull);
                    dagg_segments<KW, MODE, NV, SPEC>(P, tab, b, t, NT, lane);
                    did = true;
                }
                if (__syncthreads_or(did ? 1 : 0)) {
                    for (uint32_t i = tid; i < K; i += AT) flush_slot(tab, K, i, ((uint64_t)b << P.shift) + i, P);
                    __syncthreads();
                }
            }
            return;
        }
        for (long long b = e0 / NB; b * NB < e1 && b < P.nbins; b++) {
            const long long ba = max(e0, b * NB) - b * NB, bb = min(e1, (b + 1) * NB) - b * NB;
            const long long t_lo = ba * TBLK, t_hi = min(NT, bb * TBLK);
            for (long long t = t_lo + (long long)warp * TG; t < t_hi; t += (long long)NW * TG)
                dagg_segments<KW, MODE, NV, SPEC>(P, tab, (int)b, t, t_hi, lane);
            __syncthreads();
            if (t_lo < t_hi)
                for (uint32_t i = tid; i < K; i += AT) flush_slot(tab, K, i, ((uint64_t)b << P.shift) + i, P);
            __syncthreads();
        }
    } else {
        // Units are taken from a TICKET counter, TCHUNK at a time and in bin-major order: all warps of the chip stay within
        // a few thousand units of each other, i.e. on one or two slices of the build structure, however unevenly the CTAs
        // run.  (A static round-robin lets them drift: over 128 bins x 26 rounds a 2 % speed difference puts the CTAs
        // several slices apart and the probes fall out of L2 — ncu: 133 GB of DRAM reads for 32 GB of rows, r02_k2_hash.)
        constexpr long long TCHUNK = 8;
        while (true) {
            long long u0 = 0;
            if (lane == 0) u0 = (long long)atomicAdd(P.ticket, (unsigned long long)TCHUNK);
            u0 = __shfl_sync(HK_FULL_MASK, u0, 0);
            if (u0 >= total) break;
            const long long u1 = min(total, u0 + TCHUNK);
            for (long long u = u0; u < u1; u++) {
                const long long b = u / groups_per_bin;
                dagg_segments<KW, MODE, NV, SPEC>(P, tab, (int)b, (u - b * groups_per_bin) * TG, NT, lane);
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < K; i += AT) flush_slot(tab, K, i, (uint64_t)i, P);
    }
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

enum OutKind { O_COUNT = 0, O_COPY32, O_LOW32_OF_U64, O_AVG_S64, O_AVG_U64, O_AVG_F64, O_F32_OF_F64, O_F64, O_F64_OF_S64, O_F64_OF_U64,
               O_F32_OF_FX, O_AVG_FX, O_F64_OF_FX, O_COPY64 };
struct OutSpec {
    int kind;
    const void *src; // dense accumulator
    void *dst;       // output column
    double oscale;   // O_*_FX: 2^lo, value = (double)fixed-point sum * oscale
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
            case O_COPY64: reinterpret_cast<unsigned long long *>(o.dst)[g] = reinterpret_cast<const unsigned long long *>(o.src)[r]; break;
            case O_F32_OF_FX: reinterpret_cast<float *>(o.dst)[g] = (float)((double)reinterpret_cast<const long long *>(o.src)[r] * o.oscale); break;
            case O_AVG_FX: reinterpret_cast<double *>(o.dst)[g] = (double)reinterpret_cast<const long long *>(o.src)[r] * o.oscale / (double)cnt[e]; break;
            case O_F64_OF_FX: reinterpret_cast<double *>(o.dst)[g] = (double)reinterpret_cast<const long long *>(o.src)[r] * o.oscale; break;
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
    // borrowed columns (hark_table_from_device) can change under the table: their statistics are never kept
    const bool cacheable = col.owned && dtype == col.dtype && ctx->opt("stats.cache", 1) != 0;
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

// f32 zone map (see hk_f32_fxstats_kernel); cached on owned table columns like min / max
struct FxStats {
    bool finite = false, any = false;
    int hi = 0, lo = 0;
};

static int hk_f32_fxstats(hark_ctx *ctx, const void *ptr, const hark_col *col, int64_t n, FxStats *st) {
    const bool cacheable = col && col->owned && col->ptr == ptr && ctx->opt("stats.cache", 1) != 0;
    if (cacheable && col->fx_valid) {
        *st = FxStats{col->fx_finite, col->fx_any, col->fx_hi, col->fx_lo};
        return HARK_OK;
    }
    unsigned long long *d = nullptr;
    HK_TRY(ctx->dalloc((void **)&d, 4 * sizeof(unsigned long long)));
    ctx->h_scalars[0] = 0;
    ctx->h_scalars[1] = ~0ull;
    ctx->h_scalars[2] = 0;
    ctx->h_scalars[3] = 0;
    cudaError_t e = cudaMemcpyAsync(d, ctx->h_scalars, 4 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && n > 0) {
        hk_f32_fxstats_kernel<<<grid_for(ctx, n, 8), 256, 0, ctx->stream>>>((const uint32_t *)ptr, n, d);
        e = cudaGetLastError();
        ctx->count_launch();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_scalars, d, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ctx->dfree(d);
    if (e != cudaSuccess) return ctx->fail(HARK_ERR_CUDA, std::string("f32 statistics: ") + cudaGetErrorString(e));
    st->any = ctx->h_scalars[3] != 0;
    st->finite = ctx->h_scalars[2] == 0;
    st->hi = st->any ? (int)ctx->h_scalars[0] - 1024 : 0;
    st->lo = st->any ? (int)ctx->h_scalars[1] - 1024 : 0;
    if (cacheable) {
        col->fx_valid = true;
        col->fx_finite = st->finite;
        col->fx_any = st->any;
        col->fx_hi = st->hi;
        col->fx_lo = st->lo;
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
        if ((rq.agg_code[j] == HARK_AGG_AVG || rq.agg_code[j] == HARK_AGG_SUMF64 || rq.agg_code[j] == HARK_AGG_SUM64) && rq.agg_val[j] >= 0)
            col_needs_avg[rq.agg_val[j]] = true; // these need the exact 64-bit sum
    // f32 SUM / AVG: exact fixed point when the column's zone map allows it (every value a multiple of 2^lo, below
    // 2^(hi+1)): v / 2^lo is an integer of B = hi + 1 - lo bits.  B <= 31 -> one 32-bit addend per row (the integer
    // path's cost); B + log2(n) <= 62 -> 64-bit addends; else (or non-finite values) the f64 compare-and-swap path.
    struct FxPlan {
        int kind = -1; // -1 none, else A_FXSUM32 / A_FXSUM64
        float fscale = 1.0f;
        double oscale = 1.0;
    };
    std::vector<FxPlan> fxp((size_t)std::max(rq.nvals, 1));
    if (ctx->opt("dense.fixed_point", 1) != 0 && !rq.pinned_u32) {
        for (int v = 0; v < rq.nvals; v++) {
            if (rq.val_dtypes[v] != HARK_F32) continue;
            bool wanted = false;
            for (int j = 0; j < rq.c; j++)
                wanted = wanted || (rq.agg_val[j] == v && (rq.agg_code[j] == HARK_AGG_SUM || rq.agg_code[j] == HARK_AGG_AVG ||
                                                           rq.agg_code[j] == HARK_AGG_SUMF64 || rq.agg_code[j] == HARK_AGG_SUM64));
            if (!wanted) continue;
            FxStats st;
            HK_TRY(hk_f32_fxstats(ctx, rq.vals[v], rq.val_cols[v], n, &st));
            if (!st.finite) continue;
            if (!st.any) st.hi = st.lo = 0;
            const int B = st.hi + 1 - st.lo;
            int lg = 0;
            while ((1ll << lg) < n) lg++;
            if (st.lo < -126 || st.lo > 126) continue;
            if (B <= 31) fxp[v].kind = A_FXSUM32;
            else if (B + lg <= 62) fxp[v].kind = A_FXSUM64;
            else continue;
            fxp[v].fscale = ldexpf(1.0f, -st.lo);
            fxp[v].oscale = ldexp(1.0, st.lo);
        }
    }
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
            if (is_f && fxp[vi].kind >= 0) outs.push_back({O_F32_OF_FX, acc_index(vi, fxp[vi].kind, 2), HARK_F32});
            else if (is_f) outs.push_back({O_F32_OF_F64, acc_index(vi, A_FSUM, 2), HARK_F32});
            else if (col_needs_avg[vi]) outs.push_back({O_LOW32_OF_U64, acc_index(vi, is_s ? A_SUM64S : A_SUM64U, 2), vdt});
            else outs.push_back({O_COPY32, acc_index(vi, A_SUM32, 1), vdt});
            break;
        case HARK_AGG_AVG:
            if (is_f && fxp[vi].kind >= 0) outs.push_back({O_AVG_FX, acc_index(vi, fxp[vi].kind, 2), HARK_F64});
            else if (is_f) outs.push_back({O_AVG_F64, acc_index(vi, A_FSUM, 2), HARK_F64});
            else outs.push_back({is_s ? O_AVG_S64 : O_AVG_U64, acc_index(vi, is_s ? A_SUM64S : A_SUM64U, 2), HARK_F64});
            break;
        case HARK_AGG_SUM64:
            if (!is_f) {
                outs.push_back({O_COPY64, acc_index(vi, is_s ? A_SUM64S : A_SUM64U, 2), HARK_I64});
                break;
            }
            // float columns: SUM64 = SUMF64
        case HARK_AGG_SUMF64:
            if (is_f && fxp[vi].kind >= 0) outs.push_back({O_F64_OF_FX, acc_index(vi, fxp[vi].kind, 2), HARK_F64});
            else if (is_f) outs.push_back({O_F64, acc_index(vi, A_FSUM, 2), HARK_F64});
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
    const bool hash_mode = rq.htab != nullptr;
    const bool lut_mode = rq.lut != nullptr || hash_mode; // the slot comes from a lookup: one table of R <= K slots
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
    const int64_t slice_bytes = ctx->opt("join.lut_slice_bytes", 16ll << 20);
    if (hash_mode) {
        // slices of the hash table: entries [s << lut_shift, (s + 1) << lut_shift)
        const int esz = kw == 4 ? 8 : 16;
        const uint64_t tab_bytes = (rq.hmask + 1) * (uint64_t)esz;
        if ((int64_t)tab_bytes > slice_bytes * 3 / 2 && rq.nvals <= PMAXV) {
            lut_shift = floor_log2_u64((uint64_t)slice_bytes / esz);
            while ((rq.hmask >> lut_shift) + 1 > 256) lut_shift++;
            lut_bins = (int)((rq.hmask >> lut_shift) + 1);
        }
    } else if (lut_mode) {
        const uint64_t lut_bytes = (uint64_t)rq.pk_span * 4ull;
        if ((int64_t)lut_bytes > slice_bytes * 3 / 2 && rq.nvals <= PMAXV) {
            lut_shift = floor_log2_u64((uint64_t)slice_bytes / 4);
            while ((((uint64_t)rq.pk_span - 1) >> lut_shift) + 1 > 256) lut_shift++;
            lut_bins = (int)((((uint64_t)rq.pk_span - 1) >> lut_shift) + 1);
        }
    }
    const bool use_tiles = (nbins > 1 || lut_bins > 1) && (hash_mode || ctx->opt("dense.part_impl", 0) == 0);
    hk_tpart tp;
    if (use_tiles) {
        hk_part_spec ps;
        ps.dtype = rq.key_dtype;
        if (hash_mode) {
            ps.base = 0;
            ps.span = 0;
            ps.shift = lut_shift;
            ps.nbins = lut_bins;
            agg_nbins = lut_bins;
        } else if (lut_mode) {
            ps.base = kw == 4 ? (uint64_t)(rq.key_dtype == HARK_U32 ? (uint32_t)rq.pk_min : ((uint32_t)(int32_t)rq.pk_min ^ 0x80000000u))
                              : ((uint64_t)rq.pk_min ^ 0x8000000000000000ull);
            ps.span = (uint64_t)rq.pk_span;
            ps.shift = lut_shift;
            ps.nbins = lut_bins;
            agg_nbins = lut_bins;
        } else {
            ps.base = rq.g_lo;
            ps.span = R;
            ps.shift = shift;
            ps.nbins = nbins;
        }
        HK_TRY(hk_tile_partition(ctx, n, key, kw, ps, hash_mode ? rq.hmask : 0ull, rq.nvals, vals, &tp));
        scratch.adopt(tp.rows);
        scratch.adopt(tp.dir);
    } else if (nbins > 1 || lut_bins > 1) {
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
    P.htab = rq.htab;
    P.hmask = rq.hmask;
    P.t_rows = tp.rows;
    P.t_dir = tp.dir;
    P.t_num_tiles = tp.num_tiles;
    if (use_tiles && lut_mode) {
        HK_TRY(scratch.alloc((void **)&P.ticket, sizeof(unsigned long long)));
        HK_CUDA(ctx, cudaMemsetAsync(P.ticket, 0, sizeof(unsigned long long), ctx->stream));
    }
    const bool dynamic = ctx->opt("dense.dynamic", 1) != 0;
    if (use_tiles && !lut_mode && dynamic && ctx->opt("dense.home_by_rows", 0) == 0) {
        // dealt units need no row statistics: a CTA's first bin is its share of the BINS (the ticket counters do the
        // balancing), which saves the (bin, block) row prefix — two launches, 0.13 ms of config 3's 5.6
        P.t_nblk = (tp.num_tiles + TBLK - 1) / TBLK;
        HK_TRY(scratch.alloc((void **)&P.bin_ticket, sizeof(unsigned long long) * (size_t)nbins));
        HK_CUDA(ctx, cudaMemsetAsync(P.bin_ticket, 0, sizeof(unsigned long long) * (size_t)nbins, ctx->stream));
    } else if (use_tiles && !lut_mode) {
        const long long nblk = (tp.num_tiles + TBLK - 1) / TBLK, nent = nblk * nbins;
        uint32_t *blk_rows = nullptr;
        unsigned long long *cum = nullptr;
        HK_TRY(scratch.alloc((void **)&blk_rows, sizeof(uint32_t) * (size_t)nent));
        HK_TRY(scratch.alloc((void **)&cum, sizeof(unsigned long long) * (size_t)(nent + 1)));
        hk_tile_block_rows_kernel<<<grid_for(ctx, nent), 256, 0, ctx->stream>>>(tp.dir, tp.num_tiles, nbins, nblk, blk_rows);
        HK_CHECK_LAUNCH(ctx);
        hk_dense_scan_kernel<<<1, 1024, 0, ctx->stream>>>(blk_rows, cum, nent);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch(2);
        P.t_cum = cum;
        P.t_nblk = nblk;
        if (dynamic) {
            HK_TRY(scratch.alloc((void **)&P.bin_ticket, sizeof(unsigned long long) * (size_t)nbins));
            HK_CUDA(ctx, cudaMemsetAsync(P.bin_ticket, 0, sizeof(unsigned long long) * (size_t)nbins, ctx->stream));
        }
    }
    P.nvals = rq.nvals;
    for (int v = 0; v < rq.nvals; v++) P.vals[v] = (const uint32_t *)vals[v];
    P.nacc = (int)accs.size();
    P.gcnt = gcnt;
    std::vector<void *> gacc(accs.size(), nullptr);
    for (size_t i = 0; i < accs.size(); i++) {
        const int kind = accs[i].kind;
        const bool wide = kind == A_SUM64S || kind == A_SUM64U || kind == A_FSUM || kind == A_FPROD || kind == A_FXSUM32 || kind == A_FXSUM64;
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
        P.acc[i] = DAcc{accs[i].vcol, kind, word_of[i], gacc[i], fxp[(size_t)accs[i].vcol].fscale};
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
        // accumulator programme of the tile kernel (see SPEC_*): specialised when there is at most one accumulator, on
        // value column 0, sitting at table word 1
        int spec = SPEC_GENERIC;
        if (accs.empty()) spec = SPEC_COUNT;
        else if (accs.size() == 1 && accs[0].vcol == 0 && word_of[0] == 1 && rq.nvals == 1 && ctx->opt("dense.spec", 1) != 0)
            spec = accs[0].kind == A_SUM64S ? SPEC_SUM64S : accs[0].kind == A_SUM32 ? SPEC_SUM32 : accs[0].kind == A_FXSUM32 ? SPEC_FX32 : SPEC_GENERIC;
#define HK_DAGG_TILES1(KWv, MODEv)                                                                      \
    (spec == SPEC_SUM64S ? hk_dagg_tiles_kernel<KWv, MODEv, 1, SPEC_SUM64S>                            \
     : spec == SPEC_SUM32 ? hk_dagg_tiles_kernel<KWv, MODEv, 1, SPEC_SUM32>                            \
     : spec == SPEC_FX32 ? hk_dagg_tiles_kernel<KWv, MODEv, 1, SPEC_FX32> : hk_dagg_tiles_kernel<KWv, MODEv, 1, SPEC_GENERIC>)
#define HK_DAGG_PICK(KWv, MODEv)                                                                       \
    switch (rq.nvals) {                                                                                \
    case 0: kern = use_tiles ? (spec == SPEC_COUNT ? hk_dagg_tiles_kernel<KWv, MODEv, 0, SPEC_COUNT>   \
                                                   : hk_dagg_tiles_kernel<KWv, MODEv, 0, SPEC_GENERIC>) \
                             : hk_dagg_kernel<KWv, MODEv, 0>; break;                                   \
    case 1: kern = use_tiles ? HK_DAGG_TILES1(KWv, MODEv) : hk_dagg_kernel<KWv, MODEv, 1>; break;      \
    case 2: kern = use_tiles ? hk_dagg_tiles_kernel<KWv, MODEv, 2, SPEC_GENERIC> : hk_dagg_kernel<KWv, MODEv, 2>; break; \
    case 3: kern = use_tiles ? hk_dagg_tiles_kernel<KWv, MODEv, 3, SPEC_GENERIC> : hk_dagg_kernel<KWv, MODEv, 3>; break; \
    default: kern = hk_dagg_kernel<KWv, MODEv, 4>; break;                                              \
    }
        if (hash_mode) {
            if (kw == 4) { HK_DAGG_PICK(4, 2) } else { HK_DAGG_PICK(8, 2) }
        } else if (lut_mode) {
            if (kw == 4) { HK_DAGG_PICK(4, 1) } else { HK_DAGG_PICK(8, 1) }
        } else {
            if (kw == 4) { HK_DAGG_PICK(4, 0) } else { HK_DAGG_PICK(8, 0) }
        }
#undef HK_DAGG_PICK
#undef HK_DAGG_TILES1
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
        for (int j = 0; j < rq.c; j++)
            C.o[j] = OutSpec{outs[j].okind, outs[j].acc >= 0 ? gacc[outs[j].acc] : nullptr, t->cols[1 + j].ptr,
                             outs[j].acc >= 0 ? fxp[(size_t)accs[(size_t)outs[j].acc].vcol].oscale : 1.0};
        hk_dense_compact_kernel<<<(unsigned)nblk, CT, 0, ctx->stream>>>(C);
        cudaError_t e = cudaGetLastError();
        ctx->count_launch();
        if (e != cudaSuccess) {
            hk_table_free(ctx, t);
            return ctx->fail(HARK_ERR_CUDA, std::string("dense compact: ") + cudaGetErrorString(e));
        }
    }
    *out = t;
    return HARK_OK;
}
