// sort.cu — K3: CUB-free, stable LSD radix sort over SoA arrays (one read + one write of the carried arrays per
// 8-bit digit), K3t: passes over the top bits only + in-place tie repair when the composite key is much wider than
// log2(n), plus hash partitioning (one pass with digit = mix(key) % nparts).
//
// Reference: the radix sort inside GROUP BY and JOIN — futhark/groupby.fut:8-22 and join.fut:9-23 —
// is 32 stable 1-bit passes, each two scans + a full copy + a scatter of WHOLE rows.  Only its result
// (rows ascending by unsigned key, stable) is the contract.  Here:
//   * keys are normalised per column to  t = ordkey(x) - min  (or max - ordkey(x) for DESC), where
//     ordkey flips the sign bit of signed ints / IEEE-flips floats (NaN last); only ceil(bits(max-min)/8)
//     digits are sorted, so 20-bit keys cost 3 passes, not 32;
//   * a pass (default, "chunked"): per-chunk digit histogram -> one-CTA scan -> stable scatter with running offsets,
//     no inter-CTA communication; the first implementation ("onesweep": all histograms up front, chained-scan
//     look-back per (tile, digit)) is kept behind sort.impl=1 for A/B;
//   * a pass moves each carried array (key columns and payload / row ids) exactly once, through a
//     tile-local shared-memory reorder so that the global writes are contiguous runs per digit.
// Algorithmic bytes per pass = n · key width (histogram) + 2 · n · (sum of carried widths); the scatter kernel is
// issue-bound, not HBM-bound.
#include <algorithm>
#include <new>
#include <stdexcept>
#include <vector>

#include "hark_internal.cuh"
#include "sort.cuh"

namespace {

constexpr int ST = 256;          // threads per CTA (== number of digit bins)
constexpr int SI = 16;           // keys per thread
constexpr int STILE = ST * SI;   // keys per tile
constexpr int SWARPS = ST / 32;
constexpr int MAXA = HK_SORT_MAX_ARRAYS;
constexpr int MAXPASS = 8;       // digits per key column

struct DigitFn {
    int dtype;      // hark_dtype of the key column
    int desc;       // 1: descending
    int mode;       // 0: radix digit of the normalised key, 1: mix64(raw bits) % nparts
    int shift;
    uint32_t mask;
    uint32_t nparts;
    uint64_t base;  // min ordkey (asc) / max ordkey (desc)
    // integer keys in radix mode skip the generic order-key switch: order key = raw ^ xmask
    int fast;       // 0 generic, 1 integer ascending, 2 integer descending
    uint64_t xmask;
};

template <int KW> struct KeyRaw;
template <> struct KeyRaw<4> { using T = uint32_t; };
template <> struct KeyRaw<8> { using T = uint64_t; };

template <int KW>
__device__ __forceinline__ uint64_t norm_key(typename KeyRaw<KW>::T raw, const DigitFn &f) {
    uint64_t u;
    if constexpr (KW == 4) u = hk_ordkey32(raw, f.dtype);
    else u = hk_ordkey64(raw, f.dtype);
    return f.desc ? (f.base - u) : (u - f.base);
}

template <int KW>
__device__ __forceinline__ uint32_t digit_of(typename KeyRaw<KW>::T raw, const DigitFn &f) {
    if (f.fast) {
        if constexpr (KW == 4) {
            const uint32_t u = raw ^ (uint32_t)f.xmask;
            const uint32_t t = f.fast == 1 ? u - (uint32_t)f.base : (uint32_t)f.base - u;
            return (t >> f.shift) & f.mask;
        } else {
            const uint64_t u = raw ^ f.xmask;
            const uint64_t t = f.fast == 1 ? u - f.base : f.base - u;
            return (uint32_t)(t >> f.shift) & f.mask;
        }
    }
    if (f.mode == 1) return (uint32_t)(hk_mix64(0x68617368ull, 0, (uint64_t)raw) % f.nparts);
    return (uint32_t)(norm_key<KW>(raw, f) >> f.shift) & f.mask;
}

// ---- min / max of the order key of one column ----
template <int KW>
__global__ void __launch_bounds__(256) hk_minmax_kernel(const void *__restrict__ col, int64_t n, int dtype,
                                                         unsigned long long *out /* [0]=min [1]=max */) {
    using T = typename KeyRaw<KW>::T;
    const T *p = reinterpret_cast<const T *>(col);
    uint64_t lo = ~0ull, hi = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t u;
        if constexpr (KW == 4) u = hk_ordkey32(p[i], dtype);
        else u = hk_ordkey64(p[i], dtype);
        lo = min(lo, u);
        hi = max(hi, u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(HK_FULL_MASK, lo, o));
        hi = max(hi, __shfl_xor_sync(HK_FULL_MASK, hi, o));
    }
    __shared__ uint64_t slo[8], shi[8];
    if ((threadIdx.x & 31) == 0) {
        slo[threadIdx.x >> 5] = lo;
        shi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) {
            lo = min(lo, slo[w]);
            hi = max(hi, shi[w]);
        }
        atomicMin(out, (unsigned long long)lo);
        atomicMax(out + 1, (unsigned long long)hi);
    }
}

// ---- digit histograms of all passes of one key column (one read of the column) ----
struct HistParams {
    const void *key;
    int64_t n;
    int npass;
    DigitFn f[MAXPASS];
    unsigned long long *hist; // [npass][256]
};

template <int KW>
__global__ void __launch_bounds__(256) hk_hist_kernel(const __grid_constant__ HistParams P) {
    using T = typename KeyRaw<KW>::T;
    __shared__ uint32_t sh[MAXPASS][256];
    for (int i = threadIdx.x; i < MAXPASS * 256; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    const T *p = reinterpret_cast<const T *>(P.key);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        const T raw = p[i];
        if (P.f[0].mode == 1) {
            atomicAdd(&sh[0][digit_of<KW>(raw, P.f[0])], 1u);
        } else {
            const uint64_t t = norm_key<KW>(raw, P.f[0]);
            for (int q = 0; q < P.npass; q++) atomicAdd(&sh[q][(uint32_t)(t >> P.f[q].shift) & P.f[q].mask], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P.npass * 256; i += 256) {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(&P.hist[i], (unsigned long long)c);
    }
}

// exclusive scan of each pass's 256 bins (one block per pass, in place)
__global__ void __launch_bounds__(256) hk_hist_scan_kernel(unsigned long long *hist) {
    unsigned long long *h = hist + (size_t)blockIdx.x * 256;
    __shared__ unsigned long long wtot[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long c = h[threadIdx.x];
    unsigned long long inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(HK_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    unsigned long long off = 0;
    for (int w = 0; w < warp; w++) off += wtot[w];
    h[threadIdx.x] = off + inc - c;
}

// ---- one onesweep pass ----
// status word: tag(8) | flag(2: 1 aggregate, 2 inclusive) | value(54).  The tag is the pass number, so the
// status array is zeroed once per sort, not once per pass.
constexpr uint64_t SW_VAL = (1ull << 54) - 1;
__device__ __forceinline__ uint64_t sw_make(uint32_t tag, uint32_t flag, uint64_t v) {
    return ((uint64_t)tag << 56) | ((uint64_t)flag << 54) | v;
}

struct PassParams {
    DigitFn f;
    int na;      // carried arrays; array `ka` is the key column of this pass
    int ka;
    const void *in[MAXA];
    void *out[MAXA];
    int width[MAXA];
    int64_t n;
    int64_t num_tiles;
    uint64_t *status; // [num_tiles][256]
    unsigned long long *ticket;
    const unsigned long long *pass_offsets; // [256]
    uint32_t tag;
};

template <int KW>
__global__ void __launch_bounds__(ST, 2) hk_onesweep_kernel(const __grid_constant__ PassParams P) {
    using KT = typename KeyRaw<KW>::T;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    uint64_t *stage = reinterpret_cast<uint64_t *>(s_dyn); // STILE slots of 8 bytes
    __shared__ uint32_t wh[SWARPS][256];
    __shared__ uint32_t s_binstart[256];
    __shared__ uint64_t s_gbase[256];
    __shared__ uint8_t s_digit[STILE];
    __shared__ uint32_t s_wtot[SWARPS];
    __shared__ long long s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const KT *keyp = reinterpret_cast<const KT *>(P.in[P.ka]);

    while (true) {
        if (tid == 0) s_tile = (long long)atomicAdd(P.ticket, 1ull);
        for (int i = tid; i < SWARPS * 256; i += ST) (&wh[0][0])[i] = 0;
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.num_tiles) break;
        const int64_t tile_base = tile * STILE;
        const int count = (int)min((int64_t)STILE, P.n - tile_base);

        // ---- keys, warp-striped: item i of lane l in warp w is tile element w*512 + i*32 + l ----
        KT key[SI];
        uint32_t rank[SI];
#pragma unroll
        for (int i = 0; i < SI; i++) {
            const int idx = warp * (SI * 32) + i * 32 + lane;
            key[i] = idx < count ? keyp[tile_base + idx] : (KT)0;
        }
        // ---- stable rank of every key among the keys of its warp with the same digit ----
#pragma unroll
        for (int i = 0; i < SI; i++) {
            const int idx = warp * (SI * 32) + i * 32 + lane;
            const bool valid = idx < count;
            const uint32_t d = valid ? digit_of<KW>(key[i], P.f) : 256u;
            const uint32_t peers = __match_any_sync(HK_FULL_MASK, d);
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (valid && lane == leader) {
                old = wh[warp][d];
                wh[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(HK_FULL_MASK, old, leader);
            rank[i] = old + __popc(peers & lt_mask);
            __syncwarp();
        }
        __syncthreads();

        // ---- per digit bin (thread b owns bin b): warp offsets, tile offsets, chained scan ----
        {
            const int b = tid;
            uint32_t sum = 0;
#pragma unroll
            for (int w = 0; w < SWARPS; w++) {
                const uint32_t c = wh[w][b];
                wh[w][b] = sum;
                sum += c;
            }
            const uint32_t inc = hk_warp_incl_scan_u32(sum);
            if (lane == 31) s_wtot[warp] = inc;
            __syncthreads();
            uint32_t woff = 0;
            for (int w = 0; w < warp; w++) woff += s_wtot[w];
            const uint32_t binstart = woff + inc - sum;
            s_binstart[b] = binstart;

            uint64_t *my = P.status + (size_t)tile * 256 + b;
            uint64_t excl = 0;
            if (tile == 0) {
                hk_st_relaxed_u64(my, sw_make(P.tag, 2, sum));
            } else {
                hk_st_relaxed_u64(my, sw_make(P.tag, 1, sum));
                int64_t t = tile - 1;
                while (true) {
                    uint64_t v;
                    do {
                        v = hk_ld_relaxed_u64(P.status + (size_t)t * 256 + b);
                    } while ((uint32_t)(v >> 56) != P.tag || ((v >> 54) & 3) == 0);
                    excl += v & SW_VAL;
                    if (((v >> 54) & 3) == 2) break;
                    t--;
                }
                hk_st_relaxed_u64(my, sw_make(P.tag, 2, excl + sum));
            }
            s_gbase[b] = (uint64_t)P.pass_offsets[b] + excl - (uint64_t)binstart; // wraps; undone by + position
        }
        __syncthreads();

        // ---- tile-local reorder of the keys, remember every item's slot ----
#pragma unroll
        for (int i = 0; i < SI; i++) {
            const int idx = warp * (SI * 32) + i * 32 + lane;
            if (idx < count) {
                const uint32_t d = digit_of<KW>(key[i], P.f);
                const uint32_t pos = s_binstart[d] + wh[warp][d] + rank[i];
                rank[i] = pos;
                stage[pos] = (uint64_t)key[i];
                s_digit[pos] = (uint8_t)d;
            }
        }
        __syncthreads();
        {
            KT *o = reinterpret_cast<KT *>(P.out[P.ka]);
            for (int j = tid; j < count; j += ST) o[s_gbase[s_digit[j]] + (uint64_t)j] = (KT)stage[j];
        }
        // ---- the other carried arrays ride the same permutation ----
        for (int a = 0; a < P.na; a++) {
            if (a == P.ka) continue;
            __syncthreads();
            if (P.width[a] == 4) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(P.in[a]);
#pragma unroll
                for (int i = 0; i < SI; i++) {
                    const int idx = warp * (SI * 32) + i * 32 + lane;
                    if (idx < count) stage[rank[i]] = (uint64_t)src[tile_base + idx];
                }
                __syncthreads();
                uint32_t *o = reinterpret_cast<uint32_t *>(P.out[a]);
                for (int j = tid; j < count; j += ST) o[s_gbase[s_digit[j]] + (uint64_t)j] = (uint32_t)stage[j];
            } else {
                const uint64_t *src = reinterpret_cast<const uint64_t *>(P.in[a]);
#pragma unroll
                for (int i = 0; i < SI; i++) {
                    const int idx = warp * (SI * 32) + i * 32 + lane;
                    if (idx < count) stage[rank[i]] = src[tile_base + idx];
                }
                __syncthreads();
                uint64_t *o = reinterpret_cast<uint64_t *>(P.out[a]);
                for (int j = tid; j < count; j += ST) o[s_gbase[s_digit[j]] + (uint64_t)j] = stage[j];
            }
        }
        __syncthreads(); // stage / s_digit / wh are rewritten by the next tile
    }
}


// ------------------------------------------------------------------------------------------------
// K3 v2 — chunked stable pass (default).  Same contract as one onesweep pass (stable scatter of all carried arrays
// by one 8-bit digit), different inter-CTA protocol: instead of a chained-scan look-back per (tile, bin), a
// histogram kernel first counts every CHUNK's rows per bin (one extra streaming read of the pass's key column), a
// one-CTA scan turns that into each chunk's start offset per bin, and the scatter kernel walks its chunk tile by
// tile with running offsets in shared memory — no global atomics, no spinning, and the next tile's keys are loaded
// while the current tile is written out.  Stability: chunks, tiles, warps, items and lanes are all visited in
// row order (warp-striped items ranked with match.any ballots).
// ------------------------------------------------------------------------------------------------
constexpr int LT = 512;          // threads per CTA
constexpr int LI = 8;            // rows per thread
constexpr int LTILE = LT * LI;   // rows per tile (4096)
constexpr int LWARPS = LT / 32;

struct LsdParams {
    DigitFn f;
    int na;      // carried arrays; array `ka` is the key column of this pass
    int ka;
    const void *in[MAXA];
    void *out[MAXA];
    int width[MAXA];
    int64_t n;
    int64_t num_tiles;
    int64_t tiles_per_chunk;
    int num_chunks;
    uint32_t *chunk_counts;           // [num_chunks][256]
    unsigned long long *chunk_base;   // [num_chunks][256]
    unsigned long long *offsets;      // [257]
    // peer mode (K8c): digit = destination GPU; peer_out[a * HK_PEER_MAX + d] = byte address such that element
    // (chunk_base + position) of array a lands at its final slot in GPU d's receive arena (NVLink stores)
    const unsigned long long *peer_out;
};

template <int KW>
__global__ void __launch_bounds__(1024) hk_lsd_hist_kernel(const __grid_constant__ LsdParams P) {
    using T = typename KeyRaw<KW>::T;
    __shared__ uint32_t sh[256];
    if (threadIdx.x < 256) sh[threadIdx.x] = 0;
    __syncthreads();
    const T *p = reinterpret_cast<const T *>(P.in[P.ka]);
    constexpr int V = 16 / KW;
    const int64_t r0 = (int64_t)blockIdx.x * P.tiles_per_chunk * LTILE;
    const int64_t r1 = min(P.n, r0 + P.tiles_per_chunk * LTILE);
    const int64_t nvec = r1 > r0 ? (r1 - r0) / V : 0;
    const T *q = p + r0;
#pragma unroll 4
    for (int64_t i = threadIdx.x; i < nvec; i += 1024) {
        T x[V];
        if constexpr (KW == 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(q) + i);
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        } else {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(q) + i);
            x[0] = v.x; x[1] = v.y;
        }
#pragma unroll
        for (int e = 0; e < V; e++) atomicAdd(&sh[digit_of<KW>(x[e], P.f)], 1u);
    }
    if (r1 > r0 && threadIdx.x < (int)((r1 - r0) - nvec * V)) atomicAdd(&sh[digit_of<KW>(q[nvec * V + threadIdx.x], P.f)], 1u);
    __syncthreads();
    if (threadIdx.x < 256) P.chunk_counts[(size_t)blockIdx.x * 256 + threadIdx.x] = sh[threadIdx.x];
}

// one CTA of 1024 threads: thread (q, b) scans quarter q of the chunks of bin b
__global__ void __launch_bounds__(1024) hk_lsd_scan_kernel(const __grid_constant__ LsdParams P) {
    __shared__ unsigned long long s_q[4][256];
    __shared__ unsigned long long wtot[8];
    const int b = threadIdx.x & 255, q = threadIdx.x >> 8;
    const int per = (P.num_chunks + 3) / 4;
    const int c0 = min(P.num_chunks, q * per), c1 = min(P.num_chunks, c0 + per);
    unsigned long long sum = 0;
#pragma unroll 8
    for (int c = c0; c < c1; c++) sum += P.chunk_counts[(size_t)c * 256 + b];
    s_q[q][b] = sum;
    __syncthreads();
    const unsigned long long total = s_q[0][b] + s_q[1][b] + s_q[2][b] + s_q[3][b];
    const int lane = b & 31, warp = b >> 5;
    unsigned long long inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(HK_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
    }
    if (q == 0 && lane == 31) wtot[warp] = inc;
    __syncthreads();
    if (q == 0) {
        unsigned long long off = 0;
        for (int w = 0; w < warp; w++) off += wtot[w];
        const unsigned long long st = off + inc - total;
        P.offsets[b] = st;
        if (b == 255) P.offsets[256] = off + inc;
        const unsigned long long q0 = sum, q1 = s_q[1][b], q2 = s_q[2][b];
        s_q[0][b] = st;
        s_q[1][b] = st + q0;
        s_q[2][b] = st + q0 + q1;
        s_q[3][b] = st + q0 + q1 + q2;
    }
    __syncthreads();
    unsigned long long run = s_q[q][b];
#pragma unroll 8
    for (int c = c0; c < c1; c++) {
        P.chunk_base[(size_t)c * 256 + b] = run;
        run += P.chunk_counts[(size_t)c * 256 + b];
    }
}

template <int KW>
__device__ __forceinline__ void lsd_load_keys(const typename KeyRaw<KW>::T *keyp, int64_t tile_base, int count, int warp, int lane,
                                              typename KeyRaw<KW>::T (&key)[LI]) {
    using KT = typename KeyRaw<KW>::T;
#pragma unroll
    for (int i = 0; i < LI; i++) {
        const int idx = warp * (LI * 32) + i * 32 + lane;
        key[i] = idx < count ? keyp[tile_base + idx] : (KT)0;
    }
}

// values of one carried array for this thread's items (warp-striped), widened to 64 bits
__device__ __forceinline__ void lsd_load_vals(const void *src, int width, int64_t tile_base, int count, int warp, int lane,
                                              uint64_t (&v)[LI]) {
    if (width == 4) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(src) + tile_base;
#pragma unroll
        for (int i = 0; i < LI; i++) {
            const int idx = warp * (LI * 32) + i * 32 + lane;
            v[i] = idx < count ? (uint64_t)p[idx] : 0ull;
        }
    } else {
        const uint64_t *p = reinterpret_cast<const uint64_t *>(src) + tile_base;
#pragma unroll
        for (int i = 0; i < LI; i++) {
            const int idx = warp * (LI * 32) + i * 32 + lane;
            v[i] = idx < count ? p[idx] : 0ull;
        }
    }
}

// staged tile -> global, bin runs contiguous; fixed trip count for full tiles so the shared-memory look-ups pipeline
template <typename OT>
__device__ __forceinline__ void lsd_write_out(OT *o, const uint64_t *stage, const uint8_t *s_digit, const uint64_t *s_gbase,
                                              int count, int tid) {
    if (count == LTILE) {
#pragma unroll
        for (int i = 0; i < LI; i++) {
            const int j = i * LT + tid;
            o[s_gbase[s_digit[j]] + (uint64_t)j] = (OT)stage[j];
        }
    } else {
        for (int j = tid; j < count; j += LT) o[s_gbase[s_digit[j]] + (uint64_t)j] = (OT)stage[j];
    }
}

// peer mode: the base pointer depends on the digit (= destination GPU); s_peer = this array's HK_PEER_MAX addresses
template <typename OT>
__device__ __forceinline__ void lsd_write_out_peer(const unsigned long long *s_peer, const uint64_t *stage, const uint8_t *s_digit,
                                                   const uint64_t *s_gbase, int count, int tid) {
    if (count == LTILE) {
#pragma unroll
        for (int i = 0; i < LI; i++) {
            const int j = i * LT + tid;
            const int d = s_digit[j];
            reinterpret_cast<OT *>(s_peer[d])[s_gbase[d] + (uint64_t)j] = (OT)stage[j];
        }
    } else {
        for (int j = tid; j < count; j += LT) {
            const int d = s_digit[j];
            reinterpret_cast<OT *>(s_peer[d])[s_gbase[d] + (uint64_t)j] = (OT)stage[j];
        }
    }
}

// One step of the ballot multisplit: peers &= lanes whose digit agrees with mine in the bit `bitmask`.
// peers & ~(ballot ^ m), m = all-ones when my bit is set: one LOP3 (LUT 0x90) after the vote.
__device__ __forceinline__ uint32_t hk_ballot_step(uint32_t peers, uint32_t d, uint32_t bitmask) {
    uint32_t out;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " .reg .b32 b, m, t;\n"
        " and.b32 t, %2, %3;\n"
        " setp.ne.u32 p, t, 0;\n"
        " vote.sync.ballot.b32 b, p, 0xffffffff;\n"
        " selp.b32 m, 0xffffffff, 0, p;\n"
        " lop3.b32 %0, %1, b, m, 0x90;\n"
        "}\n"
        : "=r"(out)
        : "r"(peers), "r"(d), "r"(bitmask));
    return out;
}

// RANK: 0 = match.any, 1 = eight ballots (register-only multisplit), 2 = lean ballots (default).  PEER: outputs go to per-digit base addresses
// (other GPUs' arenas) and the key array itself (the destination digit) is not written anywhere.
// FUSE2 (two carried arrays, the common shape: key + payload, or two key columns): both arrays are reordered in ONE
// staging round — one barrier pair and one (digit, base) look-up per row instead of two — and both arrays of the next
// tile are requested before the write-out (round 2; the per-array rounds cost the kernel most of its barrier stalls).
template <int KW, int RANK, bool PEER = false, bool FUSE2 = false>
__global__ void __launch_bounds__(LT, 2) hk_lsd_scatter_kernel(const __grid_constant__ LsdParams P) {
    using KT = typename KeyRaw<KW>::T;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    uint64_t *stage = reinterpret_cast<uint64_t *>(s_dyn); // LTILE slots of 8 bytes
    uint64_t *stage2 = stage + LTILE;                      // FUSE2: the other array's slots
    __shared__ uint32_t wh[LWARPS][256];
    __shared__ uint32_t s_binstart[256];
    __shared__ uint64_t s_gbase[256];
    __shared__ uint64_t s_run[256];
    __shared__ uint8_t s_digit[LTILE];
    __shared__ uint32_t s_wtot[8];
    __shared__ unsigned long long s_peer[PEER ? MAXA * HK_PEER_MAX : 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const KT *keyp = reinterpret_cast<const KT *>(P.in[P.ka]);
    const int64_t t0 = (int64_t)blockIdx.x * P.tiles_per_chunk;
    const int64_t t1 = min(P.num_tiles, t0 + P.tiles_per_chunk);
    if (tid < 256) s_run[tid] = P.chunk_base[(size_t)blockIdx.x * 256 + tid];
    for (int i = tid; i < LWARPS * 256; i += LT) (&wh[0][0])[i] = 0;
    if constexpr (PEER)
        for (int i = tid; i < P.na * HK_PEER_MAX; i += LT) s_peer[i] = P.peer_out[i];
    if (t0 >= t1) return;
    const int first_other = P.ka == 0 ? (P.na > 1 ? 1 : -1) : 0;

    KT key[LI];
    int count = (int)min((int64_t)LTILE, P.n - t0 * LTILE);
    lsd_load_keys<KW>(keyp, t0 * LTILE, count, warp, lane, key);
    __syncthreads();

    uint64_t v[LI];
    if constexpr (FUSE2) lsd_load_vals(P.in[first_other], P.width[first_other], t0 * LTILE, count, warp, lane, v);
    for (int64_t tile = t0; tile < t1; tile++) {
        const int64_t tile_base = tile * LTILE;
        const int cur_count = count;
        // the first carried array's values are requested now and arrive while the keys are being ranked
        if constexpr (!FUSE2)
            if (first_other >= 0) lsd_load_vals(P.in[first_other], P.width[first_other], tile_base, cur_count, warp, lane, v);
        // ---- stable rank of every key among the keys of its warp with the same digit ----
        // (issuing the 8 counter updates as back-to-back atomics and reading the results after the loop was tried:
        //  the extra live registers spill under the 64-register cap and the pass got 7 % slower)
        uint32_t rank[LI];
        if constexpr (RANK == 2) {
            // lean variant of the ballot ranking (ncu: the ranking loop was 55 % of the kernel's instructions): digits
            // are extracted once with the pass constants in registers, a ballot step is 4 SASS instructions (bit test,
            // vote, select, one LOP3), the highest peer lane leads (FLO, no BREV), full tiles skip the validity votes
            uint32_t dpk[LI / 4];
            if (P.f.fast == 1) {
                const KT xm = (KT)P.f.xmask, base = (KT)P.f.base;
                const int sh = P.f.shift;
                const uint32_t mk = P.f.mask;
#pragma unroll
                for (int i = 0; i < LI; i++) {
                    const uint32_t d = (uint32_t)(((key[i] ^ xm) - base) >> sh) & mk;
                    if ((i & 3) == 0) dpk[i >> 2] = d;
                    else dpk[i >> 2] |= d << ((i & 3) * 8);
                }
            } else {
#pragma unroll
                for (int i = 0; i < LI; i++) {
                    const uint32_t d = digit_of<KW>(key[i], P.f) & 0xffu;
                    if ((i & 3) == 0) dpk[i >> 2] = d;
                    else dpk[i >> 2] |= d << ((i & 3) * 8);
                }
            }
            const bool full = cur_count == LTILE;
            uint32_t *whw = &wh[warp][0];
#pragma unroll
            for (int i = 0; i < LI; i++) {
                const int idx = warp * (LI * 32) + i * 32 + lane;
                const uint32_t d = (dpk[i >> 2] >> ((i & 3) * 8)) & 0xffu;
                bool valid = true;
                uint32_t peers = 0xffffffffu;
                if (!full) {
                    valid = idx < cur_count;
                    peers = __ballot_sync(HK_FULL_MASK, valid);
                    if (!valid) peers = ~peers;
                }
#pragma unroll
                for (int k = 0; k < 8; k++) peers = hk_ballot_step(peers, d, 1u << k);
                const int leader = 31 - __clz((int)peers);
                uint32_t old = 0;
                if (valid && lane == leader) {
                    old = whw[d];
                    whw[d] = old + __popc(peers);
                }
                old = __shfl_sync(HK_FULL_MASK, old, leader);
                rank[i] = (valid ? (d << 16) : 0xffff0000u) | (old + __popc(peers & lt_mask));
                __syncwarp();
            }
        } else {
#pragma unroll
        for (int i = 0; i < LI; i++) {
            const int idx = warp * (LI * 32) + i * 32 + lane;
            const bool valid = idx < cur_count;
            const uint32_t d = valid ? digit_of<KW>(key[i], P.f) : 256u;
            uint32_t peers;
            if constexpr (RANK == 0) {
                peers = __match_any_sync(HK_FULL_MASK, d);
            } else {
                peers = __ballot_sync(HK_FULL_MASK, valid);
                if (!valid) peers = ~peers;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t b = __ballot_sync(HK_FULL_MASK, (d >> k) & 1u);
                    peers &= ((d >> k) & 1u) ? b : ~b;
                }
            }
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (valid && lane == leader) {
                old = wh[warp][d];
                wh[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(HK_FULL_MASK, old, leader);
            rank[i] = (valid ? (d << 16) : 0xffff0000u) | (old + __popc(peers & lt_mask));
            __syncwarp();
        }
        }
        __syncthreads();
        // ---- thread b owns bin b: offsets of the warps inside the bin, bin start inside the tile, global base ----
        {
            uint32_t sum = 0;
            if (tid < 256) {
#pragma unroll
                for (int w = 0; w < LWARPS; w++) {
                    const uint32_t c = wh[w][tid];
                    wh[w][tid] = sum;
                    sum += c;
                }
            }
            const uint32_t inc = hk_warp_incl_scan_u32(sum);
            if (lane == 31 && warp < 8) s_wtot[warp] = inc;
            __syncthreads();
            if (tid < 256) {
                uint32_t woff = 0;
                for (int w = 0; w < warp; w++) woff += s_wtot[w];
                const uint32_t binstart = woff + inc - sum;
                s_binstart[tid] = binstart;
                const uint64_t run = s_run[tid];
                s_gbase[tid] = run - (uint64_t)binstart; // wraps; undone by + position
                s_run[tid] = run + sum;
            }
        }
        __syncthreads();
        // ---- tile-local reorder of the keys; every item remembers its slot ----
#pragma unroll
        for (int i = 0; i < LI; i++) {
            const uint32_t d = rank[i] >> 16;
            if (d < 256u) {
                const uint32_t pos = s_binstart[d] + wh[warp][d] + (rank[i] & 0xffffu);
                rank[i] = pos;
                stage[pos] = (uint64_t)key[i];
                if constexpr (FUSE2) stage2[pos] = v[i];
                s_digit[pos] = (uint8_t)d;
            }
        }
        // the key registers are free: load the next tile's keys now, they arrive while this tile is written out
        if (tile + 1 < t1) {
            count = (int)min((int64_t)LTILE, P.n - (tile + 1) * LTILE);
            lsd_load_keys<KW>(keyp, (tile + 1) * LTILE, count, warp, lane, key);
            if constexpr (FUSE2) lsd_load_vals(P.in[first_other], P.width[first_other], (tile + 1) * LTILE, count, warp, lane, v);
        }
        __syncthreads();
        for (int i = tid; i < LWARPS * 256; i += LT) (&wh[0][0])[i] = 0; // last read above; next written after >= 1 barrier
        if constexpr (FUSE2) {
            // both arrays leave in one sweep: one (digit, base) look-up per row
            KT *ok = reinterpret_cast<KT *>(P.out[P.ka]);
            const bool w4 = P.width[first_other] == 4;
            uint32_t *o4 = reinterpret_cast<uint32_t *>(P.out[first_other]);
            uint64_t *o8 = reinterpret_cast<uint64_t *>(P.out[first_other]);
            if (cur_count == LTILE) {
#pragma unroll
                for (int i = 0; i < LI; i++) {
                    const int j = i * LT + tid;
                    const uint64_t g = s_gbase[s_digit[j]] + (uint64_t)j;
                    ok[g] = (KT)stage[j];
                    if (w4) o4[g] = (uint32_t)stage2[j];
                    else o8[g] = stage2[j];
                }
            } else {
                for (int j = tid; j < cur_count; j += LT) {
                    const uint64_t g = s_gbase[s_digit[j]] + (uint64_t)j;
                    ok[g] = (KT)stage[j];
                    if (w4) o4[g] = (uint32_t)stage2[j];
                    else o8[g] = stage2[j];
                }
            }
            __syncthreads(); // stage / stage2 / s_digit are rewritten by the next tile
            continue;
        }
        if constexpr (!PEER) lsd_write_out<KT>(reinterpret_cast<KT *>(P.out[P.ka]), stage, s_digit, s_gbase, cur_count, tid);
        // ---- the other carried arrays ride the same permutation; array a+1 is loaded while array a is written ----
        for (int a = first_other; a >= 0 && a < P.na;) {
            __syncthreads(); // everyone has read the previous array out of `stage`
#pragma unroll
            for (int i = 0; i < LI; i++) {
                const int idx = warp * (LI * 32) + i * 32 + lane;
                if (idx < cur_count) stage[rank[i]] = v[i];
            }
            int nxt = a + 1;
            if (nxt == P.ka) nxt++;
            if (nxt < P.na) lsd_load_vals(P.in[nxt], P.width[nxt], tile_base, cur_count, warp, lane, v);
            __syncthreads();
            if constexpr (PEER) {
                if (P.width[a] == 4) lsd_write_out_peer<uint32_t>(s_peer + a * HK_PEER_MAX, stage, s_digit, s_gbase, cur_count, tid);
                else lsd_write_out_peer<uint64_t>(s_peer + a * HK_PEER_MAX, stage, s_digit, s_gbase, cur_count, tid);
            } else {
                if (P.width[a] == 4) lsd_write_out<uint32_t>(reinterpret_cast<uint32_t *>(P.out[a]), stage, s_digit, s_gbase, cur_count, tid);
                else lsd_write_out<uint64_t>(reinterpret_cast<uint64_t *>(P.out[a]), stage, s_digit, s_gbase, cur_count, tid);
            }
            a = nxt;
        }
        __syncthreads(); // stage / s_digit are rewritten by the next tile
    }
}

// ------------------------------------------------------------------------------------------------
// K3b — "sweep16": the LSD pass for the 16-byte row shape (two 8-byte carried arrays: config 4's ORDER BY col1, col2).
//
// Between passes the rows live ARRAY-OF-STRUCTS ({a, b} = 16 bytes, 16-byte aligned), so after the tile-local reorder
// every digit's run in the stage is one contiguous, 16-byte aligned block whose destination is 16-byte aligned too —
// and a run leaves the SM as ONE TMA bulk copy (cp.async.bulk shared -> global, SASS UBLKCP) issued by the thread that
// owns the digit: 256 copies per 4096-row tile instead of 2 x 4096 look-up + store sequences.  A microbenchmark of
// exactly this pattern (tools/micro/bulk_small.cu, profiles/r02_micro_tma_small_bulk_copies.txt) sustains 5.9 TB/s with
// 256-byte runs.  Two protocols give a tile its global base per digit:
//   CHUNKED (default)  a chunk histogram over the pass's input (hk_sweep16_hist_kernel) + scan, then every CTA walks its
//                      contiguous chunk of tiles with running per-digit offsets; the next tile's rows are requested as
//                      soon as the row registers are free;
//   look-back (A/B)    global digit histograms of all passes up front (order-independent: one read of each key column),
//                      tiles claimed in order from a ticket, a decoupled look-back per (tile, digit) over status words
//                      (tag | flag | count).  Slower here: with ~300 tiles of 4096 rows in flight the look-back walks
//                      ~100 predecessors of 2 KB each per tile.
// The first pass reads the caller's SoA columns, the last one writes SoA again (per-row stores), so nothing outside this
// file sees the AoS form.
// Stability: chunks / tickets ascend with the input order, rows are warp-striped and ranked with ballots.
// ------------------------------------------------------------------------------------------------
struct SweepParams {
    DigitFn f;
    int kf;                 // which field of the row carries this pass's key: 0 = a, 1 = b
    // a digit that STRADDLES the two key columns: the low `lsh` bits come from (f, kf) — the top bits of the less
    // significant key —, the rest from (f2, kf2) — the low bits of the more significant one
    int two, kf2, lsh;
    DigitFn f2;
    // integer keys (FAST): digit part = (((key ^ x) + b) >> s) & m with  asc: x = xmask, b = -base;  desc: x = ~xmask,
    // b = base + 1 (base - u = ~u + base + 1): no dtype / direction / mode decisions per row
    uint64_t x1, b1, x2, b2;
    uint32_t s1, m1, s2, m2;
    const uint64_t *a_in, *b_in; // FIRST: the two SoA inputs
    const ulonglong2 *rows_in;   // !FIRST
    uint64_t *a_out, *b_out;     // LAST: SoA outputs
    ulonglong2 *rows_out;        // !LAST
    int64_t n;
    int64_t num_tiles;
    uint64_t *status;            // [num_tiles][256]                                   (look-back protocol)
    unsigned long long *ticket;
    const unsigned long long *pass_offsets; // [256] exclusive global digit offsets of this pass
    uint32_t tag;
    int64_t tiles_per_chunk;     // chunk protocol (default): CTA c owns tiles [c * tiles_per_chunk, ...)
    uint32_t *chunk_counts;      // [num_chunks][256]
    const unsigned long long *chunk_base; // [num_chunks][256] global row where the chunk's rows of each digit start
};

template <bool FAST = false>
__device__ __forceinline__ uint32_t sweep_digit(const SweepParams &P, uint64_t ra, uint64_t rb) {
    if constexpr (FAST) {
        uint32_t d = (uint32_t)((((P.kf == 0 ? ra : rb) ^ P.x1) + P.b1) >> P.s1) & P.m1;
        if (P.two) d |= ((uint32_t)((((P.kf2 == 0 ? ra : rb) ^ P.x2) + P.b2) >> P.s2) & P.m2) << P.lsh;
        return d;
    } else {
        uint32_t d = digit_of<8>(P.kf == 0 ? ra : rb, P.f);
        if (P.two) d |= digit_of<8>(P.kf2 == 0 ? ra : rb, P.f2) << P.lsh;
        return d & 0xffu;
    }
}

// chunk histogram of one pass over SoA (FIRST) or AoS rows: the counterpart of hk_lsd_hist_kernel
template <bool FIRST>
__global__ void __launch_bounds__(1024) hk_sweep16_hist_kernel(const __grid_constant__ SweepParams P) {
    __shared__ uint32_t sh[256];
    if (threadIdx.x < 256) sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t r0 = (int64_t)blockIdx.x * P.tiles_per_chunk * LTILE;
    const int64_t r1 = min(P.n, r0 + P.tiles_per_chunk * LTILE);
    if (FIRST && P.two) {
        for (int64_t i = r0 + threadIdx.x; i < r1; i += 1024) atomicAdd(&sh[sweep_digit(P, __ldg(P.a_in + i), __ldg(P.b_in + i))], 1u);
    } else if constexpr (FIRST) {
        const uint64_t *p = P.kf == 0 ? P.a_in : P.b_in;
        const int64_t nvec = r1 > r0 ? (r1 - r0) / 2 : 0; // r0 is a multiple of LTILE: 16-byte aligned
        const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(p + r0);
#pragma unroll 4
        for (int64_t i = threadIdx.x; i < nvec; i += 1024) {
            const ulonglong2 v = __ldg(q + i);
            atomicAdd(&sh[digit_of<8>(v.x, P.f) & 0xffu], 1u);
            atomicAdd(&sh[digit_of<8>(v.y, P.f) & 0xffu], 1u);
        }
        if (r1 > r0 && ((r1 - r0) & 1) && threadIdx.x == 0) atomicAdd(&sh[digit_of<8>(p[r1 - 1], P.f) & 0xffu], 1u);
    } else {
#pragma unroll 4
        for (int64_t i = r0 + threadIdx.x; i < r1; i += 1024) {
            const ulonglong2 v = __ldg(P.rows_in + i);
            atomicAdd(&sh[sweep_digit(P, v.x, v.y)], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < 256) P.chunk_counts[(size_t)blockIdx.x * 256 + threadIdx.x] = sh[threadIdx.x];
}

__device__ __forceinline__ void sw_bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)), "r"(s),
                 "r"(bytes)
                 : "memory");
}

template <bool FIRST>
__device__ __forceinline__ void sweep16_load(const SweepParams &P, int64_t tile_base, int count, int warp, int lane, uint64_t (&ra)[LI],
                                             uint64_t (&rb)[LI]) {
    const bool full = count == LTILE;
#pragma unroll
    for (int i = 0; i < LI; i++) {
        const int idx = warp * (LI * 32) + i * 32 + lane;
        ra[i] = rb[i] = 0;
        if (full || idx < count) {
            if constexpr (FIRST) {
                ra[i] = __ldcs(P.a_in + tile_base + idx);
                rb[i] = __ldcs(P.b_in + tile_base + idx);
            } else {
                const ulonglong2 r = __ldcs(P.rows_in + tile_base + idx);
                ra[i] = r.x;
                rb[i] = r.y;
            }
        }
    }
}

template <bool FIRST, bool LAST, bool CHUNKED, bool FAST>
__global__ void __launch_bounds__(LT, 2) hk_sweep16_kernel(const __grid_constant__ SweepParams P) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    ulonglong2 *stage = reinterpret_cast<ulonglong2 *>(s_dyn); // LTILE rows of 16 bytes
    __shared__ uint32_t wh[LWARPS][256];
    __shared__ uint32_t s_binstart[256];
    __shared__ uint32_t s_bincount[256];
    __shared__ uint64_t s_gbase[256];
    __shared__ uint8_t s_digit[LAST ? LTILE : 1];
    __shared__ uint32_t s_wtot[8];
    __shared__ long long s_tile;
    __shared__ uint64_t s_run[CHUNKED ? 256 : 1]; // chunk protocol: next global row of every digit for this chunk

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int i = tid; i < LWARPS * 256; i += LT) (&wh[0][0])[i] = 0;
    int64_t next_tile = (int64_t)blockIdx.x * P.tiles_per_chunk;
    const int64_t end_tile = min(P.num_tiles, next_tile + P.tiles_per_chunk);
    if constexpr (CHUNKED) {
        if (tid < 256) s_run[tid] = P.chunk_base[(size_t)blockIdx.x * 256 + tid];
        __syncthreads(); // the counters zeroed above belong to other warps (the ticket barrier orders them otherwise)
    }

    uint64_t ra[LI], rb[LI];
    bool prefetched = false;
    while (true) {
        if constexpr (!CHUNKED) {
            if (tid == 0) s_tile = (long long)atomicAdd(P.ticket, 1ull);
            __syncthreads();
        }
        const int64_t tile = CHUNKED ? next_tile++ : (int64_t)s_tile;
        if (tile >= (CHUNKED ? end_tile : P.num_tiles)) break;
        const int64_t tile_base = tile * LTILE;
        const int count = (int)min((int64_t)LTILE, P.n - tile_base);
        const bool full = count == LTILE;

        // ---- rows, warp-striped: item i of lane l in warp w is tile row w*256 + i*32 + l.  In the chunk protocol the
        // next tile is known: its rows were requested before the previous tile left (see below) ----
        if (!CHUNKED || !prefetched) sweep16_load<FIRST>(P, tile_base, count, warp, lane, ra, rb);
        // ---- digits, then the lean ballot ranking of K3 (stable rank among the warp's rows of the same digit) ----
        uint32_t dpk[LI / 4];
#pragma unroll
        for (int i = 0; i < LI; i++) {
            const uint32_t d = sweep_digit<FAST>(P, ra[i], rb[i]);
            if ((i & 3) == 0) dpk[i >> 2] = d;
            else dpk[i >> 2] |= d << ((i & 3) * 8);
        }
        uint32_t rank[LI];
        uint32_t *whw = &wh[warp][0];
#pragma unroll
        for (int i = 0; i < LI; i++) {
            const int idx = warp * (LI * 32) + i * 32 + lane;
            const uint32_t d = (dpk[i >> 2] >> ((i & 3) * 8)) & 0xffu;
            bool valid = true;
            uint32_t peers = 0xffffffffu;
            if (!full) {
                valid = idx < count;
                peers = __ballot_sync(HK_FULL_MASK, valid);
                if (!valid) peers = ~peers;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) peers = hk_ballot_step(peers, d, 1u << k);
            const int leader = 31 - __clz((int)peers);
            uint32_t old = 0;
            if (valid && lane == leader) {
                old = whw[d];
                whw[d] = old + __popc(peers);
            }
            old = __shfl_sync(HK_FULL_MASK, old, leader);
            rank[i] = (valid ? (d << 16) : 0xffff0000u) | (old + __popc(peers & lt_mask));
            __syncwarp();
        }
        // the bulk copies of this CTA's previous tile must have finished READING the stage before it is overwritten
        if (!LAST && (tid & 1) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        // ---- thread b owns digit b: warp offsets inside the bin, bin start inside the tile, look-back for the base ----
        {
            uint32_t sum = 0;
            if (tid < 256) {
#pragma unroll
                for (int w = 0; w < LWARPS; w++) {
                    const uint32_t c = wh[w][tid];
                    wh[w][tid] = sum;
                    sum += c;
                }
            }
            const uint32_t inc = hk_warp_incl_scan_u32(sum);
            if (lane == 31 && warp < 8) s_wtot[warp] = inc;
            __syncthreads();
            if (tid < 256) {
                uint32_t woff = 0;
                for (int w = 0; w < warp; w++) woff += s_wtot[w];
                const uint32_t binstart = woff + inc - sum;
                s_binstart[tid] = binstart;
                s_bincount[tid] = sum;
                if constexpr (CHUNKED) {
                    const uint64_t run = s_run[tid];
                    s_gbase[tid] = run;
                    s_run[tid] = run + sum;
                } else {
                uint64_t *my = P.status + (size_t)tile * 256 + tid;
                uint64_t excl = 0;
                if (tile == 0) {
                    hk_st_relaxed_u64(my, sw_make(P.tag, 2, sum));
                } else {
                    hk_st_relaxed_u64(my, sw_make(P.tag, 1, sum));
                    int64_t t = tile - 1;
                    while (true) {
                        uint64_t v;
                        do {
                            v = hk_ld_relaxed_u64(P.status + (size_t)t * 256 + tid);
                        } while ((uint32_t)(v >> 56) != P.tag || ((v >> 54) & 3) == 0);
                        excl += v & SW_VAL;
                        if (((v >> 54) & 3) == 2) break;
                        t--;
                    }
                    hk_st_relaxed_u64(my, sw_make(P.tag, 2, excl + sum));
                }
                s_gbase[tid] = (uint64_t)P.pass_offsets[tid] + excl; // global row of this tile's first row of digit b
                }
            }
        }
        __syncthreads();
        // ---- tile-local reorder: the row goes to its slot of the stage ----
#pragma unroll
        for (int i = 0; i < LI; i++) {
            const uint32_t d = rank[i] >> 16;
            if (d < 256u) {
                const uint32_t pos = s_binstart[d] + wh[warp][d] + (rank[i] & 0xffffu);
                stage[pos] = make_ulonglong2(ra[i], rb[i]);
                if constexpr (LAST) s_digit[pos] = (uint8_t)d;
            }
        }
        if constexpr (CHUNKED) { // the row registers are free: request the next tile now
            prefetched = next_tile < end_tile;
            if (prefetched)
                sweep16_load<FIRST>(P, next_tile * LTILE, (int)min((int64_t)LTILE, P.n - next_tile * LTILE), warp, lane, ra, rb);
        }
        // every warp zeroes ITS OWN counter row (only this warp ranks into it, and it has just read it for the last
        // time): no barrier between this tile's write-out and the next tile's ranking
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) wh[warp][i * 32 + lane] = 0;
        if constexpr (!LAST) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if constexpr (!LAST) {
            // ---- every digit's run leaves as ONE bulk copy.  The copy instruction takes uniform operands, so a warp
            // issues its lanes' copies one after the other (11 SASS instructions each): the 256 digits are spread over
            // all 16 warps (even threads), not over the first 8 ----
            if ((tid & 1) == 0) {
                const int b = tid >> 1;
                const uint32_t c = s_bincount[b];
                if (c) { // in pieces of at most 16 KB (a whole tile in one digit is a 64 KB run)
                    unsigned char *g = reinterpret_cast<unsigned char *>(P.rows_out + s_gbase[b]);
                    const unsigned char *sm = reinterpret_cast<const unsigned char *>(stage + s_binstart[b]);
                    const uint32_t bytes = c * 16u;
                    for (uint32_t o = 0; o < bytes; o += 16384u) sw_bulk_store(g + o, sm + o, min(16384u, bytes - o));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        } else {
            // ---- last pass: back to the caller's SoA columns ----
            if (full) {
#pragma unroll
                for (int i = 0; i < LI; i++) {
                    const int j = i * LT + tid;
                    const int d = s_digit[j];
                    const uint64_t g = s_gbase[d] + (uint64_t)(j - (int)s_binstart[d]);
                    const ulonglong2 r = stage[j];
                    P.a_out[g] = r.x;
                    P.b_out[g] = r.y;
                }
            } else {
                for (int j = tid; j < count; j += LT) {
                    const int d = s_digit[j];
                    const uint64_t g = s_gbase[d] + (uint64_t)(j - (int)s_binstart[d]);
                    const ulonglong2 r = stage[j];
                    P.a_out[g] = r.x;
                    P.b_out[g] = r.y;
                }
            }
            __syncthreads(); // the stage is rewritten by the next tile
        }
    }
    if (!LAST && (tid & 1) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // shared memory must outlive the copies
}

// One stable pass over all carried arrays with the chunked protocol.  d_offsets_out (optional) receives the device
// pointer of the 257 exclusive bin offsets (caller frees).
int lsd_pass(hark_ctx *ctx, LsdParams &P, int kw, unsigned long long **d_offsets_out) {
    P.num_tiles = (P.n + LTILE - 1) / LTILE;
    int occ = 1;
    const int64_t rk = ctx->opt("sort.rank", 2);
    const bool fuse2 = P.na == 2 && !P.peer_out && rk == 2 && ctx->opt("sort.fuse2", 1) != 0;
    const size_t smem = (size_t)LTILE * 8 * (fuse2 ? 2 : 1);
    void (*scatter)(const LsdParams) =
        kw == 4 ? (rk == 2 ? hk_lsd_scatter_kernel<4, 2> : rk == 1 ? hk_lsd_scatter_kernel<4, 1> : hk_lsd_scatter_kernel<4, 0>)
                : (rk == 2 ? hk_lsd_scatter_kernel<8, 2> : rk == 1 ? hk_lsd_scatter_kernel<8, 1> : hk_lsd_scatter_kernel<8, 0>);
    if (fuse2) scatter = kw == 4 ? hk_lsd_scatter_kernel<4, 2, false, true> : hk_lsd_scatter_kernel<8, 2, false, true>;
    if (P.peer_out) scatter = rk == 2 ? hk_lsd_scatter_kernel<4, 2, true> : hk_lsd_scatter_kernel<4, 1, true>; // the key is the 4-byte destination digit
    cudaError_t e = cudaFuncSetAttribute(scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, scatter, LT, smem);
    if (e != cudaSuccess) return ctx->fail(HARK_ERR_CUDA, std::string("sort(pass setup): ") + cudaGetErrorString(e));
    occ = std::max(1, occ);
    const int64_t want = ctx->opt("sort.ctas_per_sm", 0);
    if (want > 0) occ = (int)std::min<int64_t>(occ, want);
    const int64_t max_chunks = (int64_t)ctx->num_sms * occ;
    P.tiles_per_chunk = std::max<int64_t>(1, (P.num_tiles + max_chunks - 1) / max_chunks);
    P.num_chunks = (int)std::max<int64_t>(1, (P.num_tiles + P.tiles_per_chunk - 1) / P.tiles_per_chunk);
    void *cc = nullptr, *cb = nullptr, *off = nullptr;
    HK_TRY(ctx->dalloc(&cc, (size_t)P.num_chunks * 256 * sizeof(uint32_t)));
    int rc = ctx->dalloc(&cb, (size_t)P.num_chunks * 256 * sizeof(unsigned long long));
    if (rc == HARK_OK) rc = ctx->dalloc(&off, 257 * sizeof(unsigned long long));
    if (rc != HARK_OK) {
        ctx->dfree(cc);
        ctx->dfree(cb);
        return rc;
    }
    P.chunk_counts = (uint32_t *)cc;
    P.chunk_base = (unsigned long long *)cb;
    P.offsets = (unsigned long long *)off;
    if (kw == 4) hk_lsd_hist_kernel<4><<<(unsigned)P.num_chunks, 1024, 0, ctx->stream>>>(P);
    else hk_lsd_hist_kernel<8><<<(unsigned)P.num_chunks, 1024, 0, ctx->stream>>>(P);
    e = cudaGetLastError();
    if (e == cudaSuccess) {
        hk_lsd_scan_kernel<<<1, 1024, 0, ctx->stream>>>(P);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        scatter<<<(unsigned)P.num_chunks, LT, smem, ctx->stream>>>(P);
        e = cudaGetLastError();
    }
    ctx->count_launch(3);
    ctx->dfree(cc);
    ctx->dfree(cb);
    if (d_offsets_out) *d_offsets_out = (unsigned long long *)off;
    else ctx->dfree(off);
    if (e != cudaSuccess) return ctx->fail(HARK_ERR_CUDA, std::string("sort(pass): ") + cudaGetErrorString(e));
    return HARK_OK;
}

template <typename IT>
__global__ void __launch_bounds__(256) hk_iota_kernel(IT *out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (IT)i;
}

template <typename VT, typename IT>
__global__ void __launch_bounds__(256) hk_gather_kernel(VT *__restrict__ dst, const VT *__restrict__ src,
                                                         const IT *__restrict__ perm, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[perm[i]];
}

// ------------------------------------------------------------------------------------------------
// K3t — truncated LSD + tie repair.  When the composite key has many more significant bits than log2(n), the LSD
// passes over its LOW bits are almost always wasted work: after a stable sort by the top T >= log2(n) + slack bits
// only a small fraction of the rows still share their T-bit prefix with a neighbour.  hk_radix_sort then runs the
// passes of the top T bits only (config 4: 5 passes instead of 11) and repairs the rest here:
//   hk_sort_fix_find_kernel  — one streaming read of the sorted key columns; a row whose predecessor has the same
//                              prefix but a LARGER suffix is an inversion; the thread at the FIRST inversion of a
//                              run of equal prefixes records (run start, run length) in a work list;
//   hk_sort_fix_apply_kernel — one thread per recorded run: stable insertion sort of the run by the suffix, then
//                              every carried array that can differ inside a run is permuted in place.
// Runs of fully equal keys (duplicates) contain no inversion and cost nothing.  A run longer than FIX_LMAX that
// needs repair, or a work list overflow, raises a flag and the caller redoes the sort with all passes; a sample of
// the keys taken before the passes keeps that from happening on clustered data (see plan_truncation).
// Find and apply are separate launches because apply rewrites the key columns that find reads.
// ------------------------------------------------------------------------------------------------
constexpr int FIX_MAXK = 8;  // key columns the repair can compare
constexpr int FIX_LMAX = 32; // longest run of equal prefixes repaired in place
constexpr int FIX_T = 256;
constexpr int FIX_I = 4;     // rows per thread and iteration (8 measured slower: 122.6 vs 117.1 ms on configs[4], profiles/r02_w)

struct FixKey {
    const void *ptr;
    int width;
    int dtype;
    int desc;
    uint64_t base;
};

struct FixParams {
    int nk;    // keys, most significant first
    int kstar; // last key with sorted digits; its bits below `shift` and all later keys are unsorted
    int shift;
    FixKey key[FIX_MAXK];
    int na;
    void *arr[MAXA];
    int width[MAXA];
    int moves[MAXA];
    int64_t n;
    unsigned long long *work; // [cap] : run start << 8 | run length
    unsigned long long cap;
    unsigned long long *count; // [0] runs recorded, [1] flag: redo with all passes
    uint64_t fx, fb;           // integer fast path of the scan (hk_sort_fix_find_kernel MODE 1 / 2)
};

__device__ __forceinline__ uint64_t fix_key(const FixKey &k, int64_t r) {
    uint64_t u;
    if (k.width == 4) u = hk_ordkey32(reinterpret_cast<const uint32_t *>(k.ptr)[r], k.dtype);
    else u = hk_ordkey64(reinterpret_cast<const uint64_t *>(k.ptr)[r], k.dtype);
    return k.desc ? (k.base - u) : (u - k.base);
}

__device__ bool fix_same_prefix(const FixParams &P, int64_t a, int64_t b) {
    if ((fix_key(P.key[P.kstar], a) >> P.shift) != (fix_key(P.key[P.kstar], b) >> P.shift)) return false;
    for (int k = P.kstar - 1; k >= 0; k--)
        if (fix_key(P.key[k], a) != fix_key(P.key[k], b)) return false;
    return true;
}

// order of rows a and b by the keys from kstar on (meaningful when their prefixes are equal)
__device__ int fix_cmp_suffix(const FixParams &P, int64_t a, int64_t b) {
    for (int k = P.kstar; k < P.nk; k++) {
        const uint64_t x = fix_key(P.key[k], a), y = fix_key(P.key[k], b);
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}

// row i is an inversion (checked by the caller).  Returns start << 8 | length of its run when i is the run's FIRST
// inversion and the run fits FIX_LMAX; 0 when another thread owns the run; ~0 when the run is too long.
__device__ unsigned long long fix_claim_run(const FixParams &P, int64_t i) {
    int64_t s = i - 1;
    while (s > 0 && fix_same_prefix(P, s - 1, s)) {
        if (fix_cmp_suffix(P, s - 1, s) > 0) return 0ull; // an earlier inversion owns this run
        s--;
        if (i - s >= FIX_LMAX) return ~0ull;
    }
    int64_t e = i + 1;
    while (e < P.n && fix_same_prefix(P, e - 1, e)) {
        e++;
        if (e - s > FIX_LMAX) return ~0ull;
    }
    return ((unsigned long long)s << 8) | (unsigned long long)(e - s);
}

// MODE 0: generic order key of the truncated key column (floats: NaN canonicalisation, IEEE flip); 1 / 2: an 8- / 4-byte
// integer column, whose normalised key is (zext(raw) ^ fx) + fb — asc: fx = sign flip, fb = -base; desc: fx = ~sign flip,
// fb = base + 1.  The scan itself is one coalesced read of that column: a row fetches its predecessor's key from the
// neighbouring lane, and only candidates (same sorted prefix, not in order) look at the other key columns.  A claimed run
// goes straight to the work list with one global atomic (runs are rare: 0.9 M in configs[4]'s 2e9 rows), so the loop has no
// barrier and no shared memory (round 2: the per-CTA list cost three barriers per 1024 rows and, with the generic key
// function, 84 instructions per row — 12 ms of the 117 ms ORDER BY, profiles/r02_fix_find_ncu.txt).
template <int MODE>
__device__ __forceinline__ uint64_t fix_scan_key(const FixParams &P, const FixKey &ks, int64_t r) {
    if constexpr (MODE == 1) return (__ldcs(reinterpret_cast<const uint64_t *>(ks.ptr) + r) ^ P.fx) + P.fb;
    else if constexpr (MODE == 2) return ((uint64_t)__ldcs(reinterpret_cast<const uint32_t *>(ks.ptr) + r) ^ P.fx) + P.fb;
    else return fix_key(ks, r);
}

template <int MODE>
__global__ void __launch_bounds__(FIX_T) hk_sort_fix_find_kernel(const __grid_constant__ FixParams P) {
    const FixKey &ks = P.key[P.kstar];
    const int lane = threadIdx.x & 31;
    const int64_t per_iter = (int64_t)FIX_T * FIX_I;
    const int64_t iters = (P.n + per_iter - 1) / per_iter;
    for (int64_t it = blockIdx.x; it < iters; it += gridDim.x) {
        const int64_t r0 = it * per_iter + threadIdx.x;
        uint64_t ta[FIX_I], tb[FIX_I];
#pragma unroll
        for (int j = 0; j < FIX_I; j++) {
            const int64_t i = r0 + (int64_t)j * FIX_T;
            tb[j] = i < P.n ? fix_scan_key<MODE>(P, ks, i) : ~0ull;
        }
#pragma unroll
        for (int j = 0; j < FIX_I; j++) {
            const int64_t i = r0 + (int64_t)j * FIX_T;
            ta[j] = __shfl_up_sync(HK_FULL_MASK, tb[j], 1);
            if (lane == 0) ta[j] = (i >= 1 && i < P.n) ? fix_scan_key<MODE>(P, ks, i - 1) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < FIX_I; j++) {
            const int64_t i = r0 + (int64_t)j * FIX_T;
            if ((ta[j] >> P.shift) != (tb[j] >> P.shift) || ta[j] < tb[j]) continue; // other prefix, or in order
            if (!(i >= 1 && i < P.n)) continue;
            bool same = true;
            for (int k = P.kstar - 1; k >= 0 && same; k--) same = fix_key(P.key[k], i - 1) == fix_key(P.key[k], i);
            if (!same) continue;
            if (ta[j] == tb[j]) { // tie on the truncated key: the later keys decide
                int c = 0;
                for (int k = P.kstar + 1; k < P.nk && c == 0; k++) {
                    const uint64_t x = fix_key(P.key[k], i - 1), y = fix_key(P.key[k], i);
                    c = x < y ? -1 : (x > y ? 1 : 0);
                }
                if (c <= 0) continue;
            }
            const unsigned long long w = fix_claim_run(P, i);
            if (w == ~0ull) {
                atomicExch(P.count + 1, 1ull);
            } else if (w) {
                const unsigned long long slot = atomicAdd(P.count, 1ull);
                if (slot < P.cap) P.work[slot] = w;
                else atomicExch(P.count + 1, 1ull);
            }
        }
    }
}

__global__ void __launch_bounds__(128) hk_sort_fix_apply_kernel(const __grid_constant__ FixParams P) {
    if (P.count[1]) return; // the caller redoes the sort
    const unsigned long long runs = min(P.count[0], P.cap);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < runs; w += stride) {
        const unsigned long long e = P.work[w];
        const int64_t s = (int64_t)(e >> 8);
        const int len = (int)(e & 0xff);
        uint8_t idx[FIX_LMAX];
        for (int j = 0; j < len; j++) { // stable insertion sort of the run's row offsets by the suffix
            int p = j;
            while (p > 0 && fix_cmp_suffix(P, s + idx[p - 1], s + j) > 0) {
                idx[p] = idx[p - 1];
                p--;
            }
            idx[p] = (uint8_t)j;
        }
        uint64_t tmp[FIX_LMAX];
        for (int a = 0; a < P.na; a++) {
            if (!P.moves[a]) continue;
            if (P.width[a] == 4) {
                uint32_t *q = reinterpret_cast<uint32_t *>(P.arr[a]) + s;
                for (int j = 0; j < len; j++) tmp[j] = q[idx[j]];
                for (int j = 0; j < len; j++) q[j] = (uint32_t)tmp[j];
            } else {
                uint64_t *q = reinterpret_cast<uint64_t *>(P.arr[a]) + s;
                for (int j = 0; j < len; j++) tmp[j] = q[idx[j]];
                for (int j = 0; j < len; j++) q[j] = tmp[j];
            }
        }
    }
}

// evenly spaced sample of the normalised key tuples: out[k * S + s] = key k of row floor(s * n / S)
struct SampleParams {
    int nk;
    FixKey key[FIX_MAXK];
    int64_t n;
    int S;
    uint64_t *out;
};

__global__ void __launch_bounds__(256) hk_sort_sample_kernel(const __grid_constant__ SampleParams P) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.S) return;
    const int64_t r = (int64_t)(((unsigned __int128)(uint64_t)s * (uint64_t)P.n) / (uint64_t)P.S);
    for (int k = 0; k < P.nk; k++) P.out[(size_t)k * P.S + s] = fix_key(P.key[k], r);
}

unsigned grid_for(hark_ctx *ctx, int64_t n, int per_sm = 8) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * per_sm));
}

int bit_width_u64(uint64_t v) {
    int b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

struct Pass {
    int key;      // index into keys
    DigitFn f;
    int hist_slot;
};

struct KeyRange {
    uint64_t lo = 0, hi = 0; // min / max order key of the column
    int bits = 0;            // significant bits of (hi - lo)
};

struct TruncPlan {
    bool on = false;
    int kstar = 0; // last key with sorted digits
    int q = 0;     // its top q digits are sorted (0: all of it)
    int shift = 0; // its bits below `shift` are unsorted
};

// Pass list (least significant first) that sorts keys 0..kstar, key kstar by its top q 8-bit digits only (q == 0 or
// 8q >= bits: all of it).  Returns the number of passes; *low_shift = lowest sorted bit of key kstar.
int build_passes(const std::vector<hk_sort_keyspec> &keys, const std::vector<KeyRange> &ranges, int kstar, int q,
                 std::vector<Pass> &passes, int *low_shift = nullptr) {
    passes.clear();
    if (low_shift) *low_shift = 0;
    for (int k = kstar; k >= 0; k--) {
        const int bits = ranges[k].bits;
        int sh0 = 0;
        if (k == kstar && q > 0 && 8 * q < bits) sh0 = bits - 8 * q;
        if (k == kstar && low_shift) *low_shift = sh0;
        for (int sh = sh0; sh < bits; sh += 8) {
            Pass p;
            p.key = k;
            p.f.dtype = keys[k].dtype;
            p.f.desc = keys[k].desc;
            p.f.mode = 0;
            p.f.shift = sh;
            p.f.mask = (bits - sh >= 8) ? 0xffu : ((1u << (bits - sh)) - 1u);
            p.f.nparts = 0;
            p.f.base = keys[k].desc ? ranges[k].hi : ranges[k].lo;
            p.f.fast = hk_dtype_int(keys[k].dtype) ? (keys[k].desc ? 2 : 1) : 0;
            p.f.xmask = keys[k].dtype == HARK_I32 ? 0x80000000ull : keys[k].dtype == HARK_I64 ? 0x8000000000000000ull : 0ull;
            p.hist_slot = (int)passes.size();
            passes.push_back(p);
        }
    }
    return (int)passes.size();
}

FixKey make_fix_key(const hk_sort_keyspec &ks, const KeyRange &r, const void *ptr, int width) {
    FixKey f;
    f.ptr = ptr;
    f.width = width;
    f.dtype = ks.dtype;
    f.desc = ks.desc;
    f.base = ks.desc ? r.hi : r.lo;
    return f;
}

// Number of passes of the plan "keys 0..kstar, key kstar by its top q digits" (the loop structure of build_passes);
// *low_shift = lowest sorted bit of key kstar.
int count_passes(const int *bits, int kstar, int q, int *low_shift) {
    int np = 0;
    *low_shift = 0;
    for (int k = kstar; k >= 0; k--) {
        int sh0 = 0;
        if (k == kstar && q > 0 && 8 * q < bits[k]) sh0 = bits[k] - 8 * q;
        if (k == kstar) *low_shift = sh0;
        for (int sh = sh0; sh < bits[k]; sh += 8) np++;
    }
    return np;
}

// Decides how many of the top bits of the composite key the LSD passes must cover — pure host code, also reachable
// without a GPU through hark_debug_plan_truncation (tests/test_host_logic.py runs it against a Python restatement).
// T starts at ceil(log2 n) + slack (a uniform key then leaves < 2^-slack of the rows tied with a neighbour) and grows
// by one digit at a time while the sample (S normalised key tuples, row-major, evenly spaced rows) still shows prefix
// ties between DIFFERENT tuples — clustered keys (floats around one exponent, ids with a common high part) keep all
// their passes, duplicates cost nothing.
void plan_from_sample(int64_t n, int nk, const int *bits, const uint64_t *row, int S, int slack, TruncPlan *out) {
    out->on = false;
    int total_bits = 0, full_passes = 0;
    for (int k = 0; k < nk; k++) {
        total_bits += bits[k];
        full_passes += (bits[k] + 7) / 8;
    }
    int T = bit_width_u64((uint64_t)(n - 1)) + slack;
    if (T + 8 > total_bits || S < 1) return; // nothing to save
    // ---- sample sorted by the full tuple; per adjacent pair: first differing key and its highest differing bit ----
    std::vector<int> order(S);
    for (int s = 0; s < S; s++) order[s] = s;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        const uint64_t *x = &row[(size_t)a * nk], *y = &row[(size_t)b * nk];
        for (int k = 0; k < nk; k++)
            if (x[k] != y[k]) return x[k] < y[k];
        return false;
    });
    std::vector<std::pair<int, int>> diff; // (first differing key, bit width of the xor) of adjacent distinct tuples
    diff.reserve(S);
    for (int s = 1; s < S; s++) {
        const uint64_t *x = &row[(size_t)order[s - 1] * nk], *y = &row[(size_t)order[s] * nk];
        for (int k = 0; k < nk; k++)
            if (x[k] != y[k]) {
                diff.emplace_back(k, bit_width_u64(x[k] ^ y[k]));
                break;
            }
    }
    // A sample of S rows sees a prefix tie of the data only if it holds both rows: ties_sample ~ ties_data * (S/n)^2.
    // The repair stays cheaper than the passes it replaces while ties_data < n/8 (which is also the work list's size),
    // and a uniform key at the default slack shows 1/4 of that.
    const double expect = (double)S * (double)S / (8.0 * (double)n);
    const int64_t thr = (int64_t)std::max(4.0, expect);
    // ---- smallest digit-aligned prefix whose sample ties stay under the threshold ----
    for (;; T += 8) {
        int acc = 0, kstar = nk - 1, q = 0;
        for (int k = 0; k < nk; k++) {
            if (acc >= T) {
                kstar = k - 1;
                q = 0;
                break;
            }
            const int b = bits[k], qd = (T - acc + 7) / 8;
            if (8 * qd >= b) {
                acc += b;
                continue;
            }
            kstar = k;
            q = qd;
            break;
        }
        int shift = 0;
        const int np = count_passes(bits, kstar, q, &shift);
        if (np >= full_passes) return;
        int64_t ties = 0;
        for (const auto &d : diff)
            if (d.first > kstar || (d.first == kstar && d.second <= shift)) ties++;
        if (ties <= thr) {
            out->on = true;
            out->kstar = kstar;
            out->q = q;
            out->shift = shift;
            return;
        }
    }
}

// Device part of the planning: an evenly spaced sample of the normalised key tuples, downloaded and handed to
// plan_from_sample.
int plan_truncation(hark_ctx *ctx, int64_t n, const std::vector<hk_sort_keyspec> &keys, const std::vector<hk_sort_array> &arrays,
                    const std::vector<KeyRange> &ranges, int full_passes, TruncPlan *out) {
    const int nk = (int)keys.size();
    out->on = false;
    (void)full_passes;
    int bits[FIX_MAXK], total_bits = 0;
    for (int k = 0; k < nk; k++) total_bits += (bits[k] = ranges[k].bits);
    const int slack = (int)std::max<int64_t>(0, ctx->opt("sort.trunc_slack", 4));
    if (bit_width_u64((uint64_t)(n - 1)) + slack + 8 > total_bits) return HARK_OK; // nothing to save: skip the sample
    const int S = (int)std::min<int64_t>(n, 32768);
    uint64_t *d_sample = nullptr;
    HK_TRY(ctx->dalloc((void **)&d_sample, sizeof(uint64_t) * (size_t)S * nk));
    SampleParams SP;
    memset(&SP, 0, sizeof SP);
    SP.nk = nk;
    for (int k = 0; k < nk; k++) SP.key[k] = make_fix_key(keys[k], ranges[k], arrays[keys[k].array].in, arrays[keys[k].array].width);
    SP.n = n;
    SP.S = S;
    SP.out = d_sample;
    hk_sort_sample_kernel<<<(S + 255) / 256, 256, 0, ctx->stream>>>(SP);
    cudaError_t e = cudaGetLastError();
    ctx->count_launch();
    std::vector<uint64_t> col((size_t)S * nk), row((size_t)S * nk);
    if (e == cudaSuccess) e = cudaMemcpyAsync(col.data(), d_sample, sizeof(uint64_t) * col.size(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ctx->dfree(d_sample);
    if (e != cudaSuccess) return ctx->fail(HARK_ERR_CUDA, std::string("sort(sample): ") + cudaGetErrorString(e));
    for (int s = 0; s < S; s++)
        for (int k = 0; k < nk; k++) row[(size_t)s * nk + k] = col[(size_t)k * S + s];
    plan_from_sample(n, nk, bits, row.data(), S, slack, out);
    return HARK_OK;
}

// Runs the two repair kernels over the sorted buffers (buffer index fb).  *redo: a run too long to repair in place
// (or a full work list) — the caller must sort again with every pass.
int fix_truncated(hark_ctx *ctx, int64_t n, const std::vector<hk_sort_keyspec> &keys, std::vector<hk_sort_array> &arrays,
                  const std::vector<KeyRange> &ranges, const TruncPlan &tp, int fb, bool *redo, int64_t *runs) {
    const int nk = (int)keys.size(), na = (int)arrays.size();
    FixParams P;
    memset(&P, 0, sizeof P);
    P.nk = nk;
    P.kstar = tp.kstar;
    P.shift = tp.shift;
    for (int k = 0; k < nk; k++) P.key[k] = make_fix_key(keys[k], ranges[k], arrays[keys[k].array].buf[fb], arrays[keys[k].array].width);
    P.na = na;
    for (int a = 0; a < na; a++) {
        P.arr[a] = arrays[a].buf[fb];
        P.width[a] = arrays[a].width;
        P.moves[a] = 1;
    }
    for (int k = 0; k < tp.kstar; k++) P.moves[keys[k].array] = 0; // equal inside a run
    P.n = n;
    P.cap = (unsigned long long)std::max<int64_t>(1024, n / 8);
    void *work = nullptr, *count = nullptr;
    HK_TRY(ctx->dalloc(&work, sizeof(unsigned long long) * P.cap));
    int rc = ctx->dalloc(&count, 2 * sizeof(unsigned long long));
    if (rc != HARK_OK) {
        ctx->dfree(work);
        return rc;
    }
    P.work = (unsigned long long *)work;
    P.count = (unsigned long long *)count;
    cudaError_t e = cudaMemsetAsync(count, 0, 2 * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) {
        const int64_t iters = (n + FIX_T * FIX_I - 1) / (FIX_T * FIX_I);
        const unsigned g = (unsigned)std::max<int64_t>(1, std::min<int64_t>(iters, (int64_t)ctx->num_sms * 8));
        const FixKey &ks = P.key[P.kstar];
        const int mode = (hk_dtype_int(ks.dtype) && ctx->opt("sort.fix_fast", 1) != 0) ? (ks.width == 8 ? 1 : 2) : 0;
        if (mode) {
            const uint64_t sign = ks.dtype == HARK_I32 ? 0x80000000ull : ks.dtype == HARK_I64 ? 0x8000000000000000ull : 0ull;
            P.fx = ks.desc ? ~sign : sign;
            P.fb = ks.desc ? ks.base + 1ull : 0ull - ks.base;
        }
        if (mode == 1) hk_sort_fix_find_kernel<1><<<g, FIX_T, 0, ctx->stream>>>(P);
        else if (mode == 2) hk_sort_fix_find_kernel<2><<<g, FIX_T, 0, ctx->stream>>>(P);
        else hk_sort_fix_find_kernel<0><<<g, FIX_T, 0, ctx->stream>>>(P);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        hk_sort_fix_apply_kernel<<<(unsigned)ctx->num_sms * 16, 128, 0, ctx->stream>>>(P);
        e = cudaGetLastError();
    }
    ctx->count_launch(2);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_scalars, count, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ctx->dfree(work);
    ctx->dfree(count);
    if (e != cudaSuccess) return ctx->fail(HARK_ERR_CUDA, std::string("sort(repair): ") + cudaGetErrorString(e));
    *runs = (int64_t)ctx->h_scalars[0];
    *redo = ctx->h_scalars[1] != 0;
    return HARK_OK;
}

// All passes of a sort over two 8-byte carried arrays with K3b.  On success arrays[0..1].buf[(npass-1)&1] hold the
// result (SoA), like the chunked path leaves them.
// One pass of K3b: a primary digit part and, when the digit straddles the two keys, a second part shifted left by lsh.
struct SweepPass {
    DigitFn f;
    int key;
    int two = 0;
    DigitFn f2;
    int key2 = 0, lsh = 0;
};

// The LSD pass list of build_passes re-cut into 8-bit digits over the CONCATENATED bit string (sorted bits of the less
// significant key, then all bits of the more significant one): config 4's 20-bit col1 no longer wastes a pass on its top 4
// bits, so the same 5 passes cover 40 bits instead of 36 and the tie repair has 16x fewer runs to fix.  Only for two keys
// that are the two carried arrays; anything else keeps the plain list.
std::vector<SweepPass> sweep_passes(const std::vector<hk_sort_keyspec> &keys, const std::vector<Pass> &passes, bool straddle) {
    std::vector<SweepPass> out;
    auto plain = [&]() {
        out.clear();
        for (auto &p : passes) {
            SweepPass s;
            s.f = p.f;
            s.key = p.key;
            out.push_back(s);
        }
        return out;
    };
    if (!straddle || keys.size() != 2 || passes.empty()) return plain();
    // passes: first those of key 1 (lowest shift first), then those of key 0 from bit 0
    int n1 = 0;
    while (n1 < (int)passes.size() && passes[n1].key == 1) n1++;
    if (n1 == 0 || n1 == (int)passes.size()) return plain();
    for (int i = n1; i < (int)passes.size(); i++)
        if (passes[i].key != 0) return plain();
    auto width_of = [](uint32_t mask) {
        int w = 0;
        while (mask) {
            w++;
            mask >>= 1;
        }
        return w;
    };
    const DigitFn k1 = passes[0].f, k0 = passes[n1].f;
    int sh0 = k1.shift;                                                   // lowest sorted bit of key 1
    const int top1 = passes[n1 - 1].f.shift + width_of(passes[n1 - 1].f.mask); // bits of key 1
    const int bits0 = passes.back().f.shift + width_of(passes.back().f.mask);  // bits of key 0
    int B = (top1 - sh0) + bits0;
    if (B % 8 != 0) sh0 = std::max(0, sh0 - (8 - B % 8));                  // spend the spare digit capacity on more bits of key 1
    B = (top1 - sh0) + bits0;
    const int L1 = top1 - sh0;
    for (int pos = 0; pos < B; pos += 8) {
        SweepPass s;
        if (pos + 8 <= L1 || pos >= L1) {
            const bool from1 = pos < L1;
            const int sh = from1 ? sh0 + pos : pos - L1;
            const int w = std::min(8, (from1 ? top1 : bits0) - sh);
            s.f = from1 ? k1 : k0;
            s.f.shift = sh;
            s.f.mask = (1u << w) - 1u;
            s.key = from1 ? 1 : 0;
        } else { // the digit straddles the boundary
            const int w1 = L1 - pos, w2 = std::min(8 - w1, bits0);
            s.f = k1;
            s.f.shift = sh0 + pos;
            s.f.mask = (1u << w1) - 1u;
            s.key = 1;
            s.two = 1;
            s.f2 = k0;
            s.f2.shift = 0;
            s.f2.mask = (1u << w2) - 1u;
            s.key2 = 0;
            s.lsh = w1;
        }
        out.push_back(s);
    }
    return out;
}

int run_sweep16(hark_ctx *ctx, int64_t n, const std::vector<hk_sort_keyspec> &keys, std::vector<hk_sort_array> &arrays,
                const std::vector<Pass> &passes_in) {
    const int fb = ((int)passes_in.size() - 1) & 1; // where the caller expects the result (its own pass count decides)
    const bool lb0 = ctx->opt("sort.sweep16", 1) == 2;
    const std::vector<SweepPass> passes = sweep_passes(keys, passes_in, !lb0 && ctx->opt("sort.straddle", 1) != 0 &&
                                                                          keys[0].array != keys[keys.size() - 1].array);
    ctx->counters["sort.last_sweep16_passes"] = (int64_t)passes.size();
    const int npass = (int)passes.size();
    const int64_t num_tiles = (n + LTILE - 1) / LTILE;
    struct Scratch {
        hark_ctx *ctx;
        std::vector<void *> v;
        ~Scratch() {
            for (void *p : v) ctx->dfree(p);
        }
        int alloc(void **p, size_t bytes) {
            int rc = ctx->dalloc(p, bytes);
            if (rc == HARK_OK) v.push_back(*p);
            return rc;
        }
        void release_now(void *p) {
            v.erase(std::remove(v.begin(), v.end(), p), v.end());
            ctx->dfree(p);
        }
    } sc{ctx, {}};
    const bool lookback = ctx->opt("sort.sweep16", 1) == 2; // kept for A/B: look-back serialises the tiles
    unsigned long long *d_hist = nullptr;
    uint64_t *d_status = nullptr;
    const size_t status_words = (size_t)num_tiles * 256;
    const int64_t want = ctx->opt("sort.ctas_per_sm", 0);
    const int occ = want > 0 ? (int)std::min<int64_t>(2, want) : 2;
    const int64_t max_chunks = (int64_t)ctx->num_sms * occ;
    const int64_t tiles_per_chunk = lookback ? num_tiles : std::max<int64_t>(1, (num_tiles + max_chunks - 1) / max_chunks);
    const int num_chunks = lookback ? 0 : (int)((num_tiles + tiles_per_chunk - 1) / tiles_per_chunk);
    uint32_t *d_cc = nullptr;
    unsigned long long *d_cb = nullptr, *d_off = nullptr;
    if (lookback) {
        // ---- global digit histograms of every pass: one read per key column, then one exclusive scan per pass ----
        HK_TRY(sc.alloc((void **)&d_hist, sizeof(unsigned long long) * 256 * (size_t)npass));
        HK_CUDA(ctx, cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * 256 * (size_t)npass, ctx->stream));
        for (int p0 = 0; p0 < npass;) {
            int p1 = p0;
            while (p1 < npass && passes[p1].key == passes[p0].key && p1 - p0 < MAXPASS) p1++;
            HistParams H;
            memset(&H, 0, sizeof H);
            H.key = arrays[keys[passes[p0].key].array].in;
            H.n = n;
            H.npass = p1 - p0;
            for (int q = p0; q < p1; q++) H.f[q - p0] = passes[q].f;
            H.hist = d_hist + (size_t)256 * p0;
            hk_hist_kernel<8><<<grid_for(ctx, n, 4), 256, 0, ctx->stream>>>(H);
            HK_CHECK_LAUNCH(ctx);
            ctx->count_launch();
            p0 = p1;
        }
        hk_hist_scan_kernel<<<npass, 256, 0, ctx->stream>>>(d_hist);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
        // ---- look-back state (tags tell the passes apart: zeroed once) and one ticket per pass ----
        HK_TRY(sc.alloc((void **)&d_status, sizeof(uint64_t) * (status_words + (size_t)npass)));
        HK_CUDA(ctx, cudaMemsetAsync(d_status, 0, sizeof(uint64_t) * (status_words + (size_t)npass), ctx->stream));
    } else {
        HK_TRY(sc.alloc((void **)&d_cc, (size_t)num_chunks * 256 * sizeof(uint32_t)));
        HK_TRY(sc.alloc((void **)&d_cb, (size_t)num_chunks * 256 * sizeof(unsigned long long)));
        HK_TRY(sc.alloc((void **)&d_off, 257 * sizeof(unsigned long long)));
    }
    // ---- AoS ping-pong buffers: pass p < last writes rows[p & 1] ----
    void *rows[2] = {nullptr, nullptr};
    if (npass >= 2) HK_TRY(sc.alloc(&rows[0], (size_t)n * 16));
    if (npass >= 3) HK_TRY(sc.alloc(&rows[1], (size_t)n * 16));
    const size_t smem = (size_t)LTILE * 16;
#define HK_SW_K(a, b, c) {hk_sweep16_kernel<a, b, c, false>, hk_sweep16_kernel<a, b, c, true>}
    void (*kerns[2][2][2][2])(const SweepParams) = {
        {{HK_SW_K(false, false, false), HK_SW_K(false, false, true)}, {HK_SW_K(false, true, false), HK_SW_K(false, true, true)}},
        {{HK_SW_K(true, false, false), HK_SW_K(true, false, true)}, {HK_SW_K(true, true, false), HK_SW_K(true, true, true)}}};
#undef HK_SW_K
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++)
            for (int c = 0; c < 2; c++)
                for (int d = 0; d < 2; d++)
                    HK_CUDA(ctx, cudaFuncSetAttribute(kerns[a][b][c][d], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool allow_fast = ctx->opt("sort.sweep16_fast", 1) != 0;
    auto fast_part = [](const DigitFn &f, uint64_t *x, uint64_t *b, uint32_t *sh, uint32_t *m) {
        *x = f.fast == 1 ? f.xmask : ~f.xmask;
        *b = f.fast == 1 ? 0ull - f.base : f.base + 1ull;
        *sh = (uint32_t)f.shift;
        *m = f.mask;
    };
    const unsigned grid = lookback ? (unsigned)std::min<int64_t>(num_tiles, max_chunks) : (unsigned)num_chunks;
    for (int p = 0; p < npass; p++) {
        const bool first = p == 0, last = p == npass - 1;
        SweepParams S;
        memset(&S, 0, sizeof S);
        S.f = passes[p].f;
        S.kf = keys[passes[p].key].array;
        S.two = passes[p].two;
        if (S.two) {
            S.f2 = passes[p].f2;
            S.kf2 = keys[passes[p].key2].array;
            S.lsh = passes[p].lsh;
        }
        const bool fast = allow_fast && S.f.fast != 0 && S.f.mode == 0 && (!S.two || (S.f2.fast != 0 && S.f2.mode == 0));
        if (fast) {
            fast_part(S.f, &S.x1, &S.b1, &S.s1, &S.m1);
            if (S.two) {
                fast_part(S.f2, &S.x2, &S.b2, &S.s2, &S.m2);
                S.m2 &= 0xffu >> S.lsh; // the generic path masks the assembled digit with 0xff
            }
        }
        S.n = n;
        S.num_tiles = num_tiles;
        S.tiles_per_chunk = tiles_per_chunk;
        if (lookback) {
            S.status = d_status;
            S.ticket = (unsigned long long *)(d_status + status_words + p);
            S.pass_offsets = d_hist + (size_t)256 * p;
            S.tag = (uint32_t)p + 1;
        } else {
            S.chunk_counts = d_cc;
            S.chunk_base = d_cb;
        }
        if (first) {
            S.a_in = (const uint64_t *)arrays[0].in;
            S.b_in = (const uint64_t *)arrays[1].in;
        } else {
            S.rows_in = (const ulonglong2 *)rows[(p - 1) & 1];
        }
        if (last) {
            // the AoS buffer this pass does not read is dead: give it back before the SoA results are allocated
            if (npass >= 3) {
                void *dead = rows[p & 1];
                rows[p & 1] = nullptr;
                sc.release_now(dead);
            }
            for (int a = 0; a < 2; a++)
                if (!arrays[a].buf[fb]) HK_TRY(ctx->dalloc(&arrays[a].buf[fb], (size_t)n * 8));
            S.a_out = (uint64_t *)arrays[0].buf[fb];
            S.b_out = (uint64_t *)arrays[1].buf[fb];
        } else {
            S.rows_out = (ulonglong2 *)rows[p & 1];
        }
        if (!lookback) {
            // chunk histogram of this pass over its input (SoA for the first pass, AoS rows after), then the chunk bases
            if (first) hk_sweep16_hist_kernel<true><<<grid, 1024, 0, ctx->stream>>>(S);
            else hk_sweep16_hist_kernel<false><<<grid, 1024, 0, ctx->stream>>>(S);
            HK_CHECK_LAUNCH(ctx);
            LsdParams L;
            memset(&L, 0, sizeof L);
            L.num_chunks = num_chunks;
            L.chunk_counts = d_cc;
            L.chunk_base = d_cb;
            L.offsets = d_off;
            hk_lsd_scan_kernel<<<1, 1024, 0, ctx->stream>>>(L);
            HK_CHECK_LAUNCH(ctx);
            ctx->count_launch(2);
        }
        kerns[first ? 1 : 0][last ? 1 : 0][lookback ? 0 : 1][fast ? 1 : 0]<<<grid, LT, smem, ctx->stream>>>(S);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
    }
    return HARK_OK;
}

} // namespace

// K8c: one stable pass keyed on a u32 destination digit (< HK_PEER_MAX) whose outputs are peer addresses.
// cols/widths: the carried table columns; d_peer_out: device array [(1 + ncols) * HK_PEER_MAX] of byte addresses
// (row 0, the digit array's, is unused).
int hk_peer_scatter_pass(hark_ctx *ctx, int64_t n, const void *digit, const void *const *cols, const int *widths, int ncols,
                         const unsigned long long *d_peer_out) {
    if (ncols + 1 > MAXA) return ctx->fail(HARK_ERR_UNSUPPORTED, "peer scatter: too many columns");
    if (n == 0) return HARK_OK;
    LsdParams P;
    memset(&P, 0, sizeof P);
    P.f = DigitFn{HARK_U32, 0, 0, 0, 0xffu, 0, 0, 1, 0};
    P.na = ncols + 1;
    P.ka = 0;
    P.in[0] = digit;
    P.width[0] = 4;
    for (int c = 0; c < ncols; c++) {
        P.in[1 + c] = cols[c];
        P.width[1 + c] = widths[c];
    }
    P.n = n;
    P.peer_out = d_peer_out;
    return lsd_pass(ctx, P, 4, nullptr);
}

int hk_iota(hark_ctx *ctx, void *out, int64_t n, int width) {
    if (n == 0) return HARK_OK;
    if (width == 4) hk_iota_kernel<uint32_t><<<grid_for(ctx, n), 256, 0, ctx->stream>>>((uint32_t *)out, n);
    else hk_iota_kernel<uint64_t><<<grid_for(ctx, n), 256, 0, ctx->stream>>>((uint64_t *)out, n);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch();
    return HARK_OK;
}

int hk_gather(hark_ctx *ctx, void *dst, const void *src, int vwidth, const void *perm, int pwidth, int64_t n) {
    if (n == 0) return HARK_OK;
    const unsigned g = grid_for(ctx, n, 16);
    if (vwidth == 4 && pwidth == 4)
        hk_gather_kernel<uint32_t, uint32_t><<<g, 256, 0, ctx->stream>>>((uint32_t *)dst, (const uint32_t *)src, (const uint32_t *)perm, n);
    else if (vwidth == 8 && pwidth == 4)
        hk_gather_kernel<uint64_t, uint32_t><<<g, 256, 0, ctx->stream>>>((uint64_t *)dst, (const uint64_t *)src, (const uint32_t *)perm, n);
    else if (vwidth == 4 && pwidth == 8)
        hk_gather_kernel<uint32_t, uint64_t><<<g, 256, 0, ctx->stream>>>((uint32_t *)dst, (const uint32_t *)src, (const uint64_t *)perm, n);
    else
        hk_gather_kernel<uint64_t, uint64_t><<<g, 256, 0, ctx->stream>>>((uint64_t *)dst, (const uint64_t *)src, (const uint64_t *)perm, n);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch();
    return HARK_OK;
}

// Sorts (or hash-partitions) `n` rows held as SoA arrays.  keys: most significant first; every key's
// array must be in `arrays`.  On success arrays[a].result is a fresh device buffer the caller owns
// (allocated from the context pool) holding array a in the final order.
int hk_radix_sort(hark_ctx *ctx, int64_t n, const std::vector<hk_sort_keyspec> &keys, std::vector<hk_sort_array> &arrays,
                  int hash_nparts, int64_t *hash_counts, hk_sort_info *info) {
    const int na = (int)arrays.size();
    if (na > MAXA) return ctx->fail(HARK_ERR_UNSUPPORTED, "sort: too many carried arrays");
    if (info) *info = hk_sort_info{};
    for (auto &a : arrays) a.result = nullptr;

    std::vector<Pass> passes;
    unsigned long long *d_hist = nullptr; // [total passes][256]
    uint64_t *d_scratch = nullptr;        // minmax pairs
    int rc = HARK_OK;
    cudaError_t e = cudaSuccess;

    auto cleanup_fail = [&](int code, const std::string &msg) {
        for (auto &a : arrays) {
            for (int b = 0; b < 2; b++)
                if (a.buf[b]) {
                    ctx->dfree(a.buf[b]);
                    a.buf[b] = nullptr;
                }
            a.result = nullptr;
        }
        ctx->dfree(d_hist);
        ctx->dfree(d_scratch);
        return msg.empty() ? code : ctx->fail(code, msg);
    };
    for (auto &a : arrays) a.buf[0] = a.buf[1] = nullptr;

    const bool chunked = ctx->opt("sort.impl", 0) != 1;
    std::vector<KeyRange> ranges;
    TruncPlan trunc;
    int full_passes = 0;
    if (n > 0 && hash_nparts == 0 && !keys.empty()) {
        // ---- 1. range of every key column -> number of significant digits ----
        const int nk = (int)keys.size();
        rc = ctx->dalloc((void **)&d_scratch, sizeof(uint64_t) * 2 * nk);
        if (rc != HARK_OK) return cleanup_fail(rc, "");
        for (int k = 0; k < nk; k++) ctx->h_scalars[2 * k] = ~0ull, ctx->h_scalars[2 * k + 1] = 0;
        e = cudaMemcpyAsync(d_scratch, ctx->h_scalars, sizeof(uint64_t) * 2 * nk, cudaMemcpyHostToDevice, ctx->stream);
        for (int k = 0; k < nk && e == cudaSuccess; k++) {
            if (keys[k].have_range) continue;
            const hk_sort_array &ka = arrays[keys[k].array];
            const unsigned g = grid_for(ctx, n, 8);
            if (ka.width == 4)
                hk_minmax_kernel<4><<<g, 256, 0, ctx->stream>>>(ka.in, n, keys[k].dtype, (unsigned long long *)d_scratch + 2 * k);
            else
                hk_minmax_kernel<8><<<g, 256, 0, ctx->stream>>>(ka.in, n, keys[k].dtype, (unsigned long long *)d_scratch + 2 * k);
            e = cudaGetLastError();
            ctx->count_launch();
        }
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(ctx->h_scalars, d_scratch, sizeof(uint64_t) * 2 * nk, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return cleanup_fail(HARK_ERR_CUDA, std::string("sort(minmax): ") + cudaGetErrorString(e));
        // ---- 2. pass list: least significant key first, least significant digit first ----
        ranges.resize(nk);
        for (int k = 0; k < nk; k++) {
            ranges[k].lo = keys[k].have_range ? keys[k].lo : ctx->h_scalars[2 * k];
            ranges[k].hi = keys[k].have_range ? keys[k].hi : ctx->h_scalars[2 * k + 1];
            ranges[k].bits = bit_width_u64(ranges[k].hi - ranges[k].lo);
        }
        full_passes = build_passes(keys, ranges, nk - 1, 0, passes);
        // ---- 2b. K3t: sort only the top T bits when a key sample says the rest is (almost) never needed ----
        if (chunked && ctx->opt("sort.trunc", 1) != 0 && nk <= FIX_MAXK && n >= 2) {
            rc = plan_truncation(ctx, n, keys, arrays, ranges, full_passes, &trunc);
            if (rc != HARK_OK) return cleanup_fail(rc, "");
            if (trunc.on) build_passes(keys, ranges, trunc.kstar, trunc.q, passes);
        }
    } else if (n > 0 && hash_nparts > 0) {
        Pass p;
        p.key = 0;
        p.f = DigitFn{keys[0].dtype, 0, 1, 0, 0xffu, (uint32_t)hash_nparts, 0, 0, 0};
        p.hist_slot = 0;
        passes.push_back(p);
    }
    int npass = (int)passes.size();
    if (info) info->passes = npass;
    ctx->counters["sort.last_fallback"] = 0;
    ctx->counters["sort.last_fix_runs"] = 0;
    ctx->counters["sort.last_passes"] = npass;
    ctx->counters["sort.last_truncated"] = 0;

    if (npass == 0) { // nothing to move: the result is a copy of the input
        for (auto &a : arrays) {
            rc = ctx->dalloc(&a.buf[0], (size_t)std::max<int64_t>(n, 1) * a.width);
            if (rc != HARK_OK) return cleanup_fail(rc, "");
            if (n > 0) rc = hk_copy_bytes(ctx, a.buf[0], a.in, n * a.width);
            if (rc != HARK_OK) return cleanup_fail(rc, "");
        }
        for (auto &a : arrays) {
            a.result = a.buf[0];
            a.buf[0] = nullptr;
        }
        if (hash_counts)
            for (int i = 0; i < hash_nparts; i++) hash_counts[i] = 0;
        ctx->dfree(d_scratch);
        return HARK_OK;
    }

    if (chunked) {
        // ---- K3 v2: every pass = chunk histogram + scan + stable scatter (see hk_lsd_scatter_kernel) ----
        for (int attempt = 0;; attempt++) {
            std::vector<const void *> cur(na);
            for (int a = 0; a < na; a++) cur[a] = arrays[a].in;
            // K3b: two 8-byte carried arrays sort as 16-byte rows with TMA run stores and look-back (see hk_sweep16_kernel)
            const bool sweep16 = na == 2 && arrays[0].width == 8 && arrays[1].width == 8 && hash_nparts == 0 && npass <= 250 &&
                                 n >= ctx->opt("sort.sweep16_min_rows", 1 << 16) && ctx->opt("sort.sweep16", 1) != 0;
            ctx->counters["sort.last_sweep16"] = sweep16 ? 1 : 0;
            if (sweep16) {
                if (attempt == 0) ctx->kernel_begin();
                rc = run_sweep16(ctx, n, keys, arrays, passes);
                if (rc != HARK_OK) return cleanup_fail(rc, "");
            }
            for (int p = 0; p < npass && !sweep16; p++) {
                const int ob = p & 1;
                for (int a = 0; a < na && rc == HARK_OK; a++)
                    if (!arrays[a].buf[ob]) rc = ctx->dalloc(&arrays[a].buf[ob], (size_t)n * arrays[a].width);
                if (rc != HARK_OK) return cleanup_fail(rc, "");
                LsdParams P;
                memset(&P, 0, sizeof P);
                P.f = passes[p].f;
                P.na = na;
                P.ka = keys[passes[p].key].array;
                for (int a = 0; a < na; a++) {
                    P.in[a] = cur[a];
                    P.out[a] = arrays[a].buf[ob];
                    P.width[a] = arrays[a].width;
                }
                P.n = n;
                unsigned long long *d_off = nullptr;
                if (p == 0 && attempt == 0) ctx->kernel_begin();
                rc = lsd_pass(ctx, P, arrays[P.ka].width, hash_counts ? &d_off : nullptr);
                if (rc != HARK_OK) return cleanup_fail(rc, "");
                if (hash_counts) { // bucket sizes for the caller
                    e = cudaMemcpyAsync(ctx->h_scalars, d_off, sizeof(uint64_t) * 256, cudaMemcpyDeviceToHost, ctx->stream);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
                    ctx->dfree(d_off);
                    if (e != cudaSuccess) return cleanup_fail(HARK_ERR_CUDA, std::string("sort(counts): ") + cudaGetErrorString(e));
                    for (int i = 0; i < hash_nparts; i++)
                        hash_counts[i] = (int64_t)((i + 1 < 256 ? ctx->h_scalars[i + 1] : (uint64_t)n) - ctx->h_scalars[i]);
                }
                for (int a = 0; a < na; a++) cur[a] = arrays[a].buf[ob];
            }
            if (!trunc.on) break;
            // ---- K3t: repair the runs the dropped low digits would have ordered ----
            bool redo = false;
            int64_t runs = 0;
            rc = fix_truncated(ctx, n, keys, arrays, ranges, trunc, (npass - 1) & 1, &redo, &runs);
            if (rc != HARK_OK) return cleanup_fail(rc, "");
            ctx->counters["sort.last_fix_runs"] = runs;
            if (!redo) break;
            trunc.on = false; // a long run needed the low digits after all: all passes, from the input
            ctx->counters["sort.last_fallback"] = 1;
            npass = build_passes(keys, ranges, (int)keys.size() - 1, 0, passes);
        }
        ctx->kernel_end();
        if (info) info->passes = npass;
        ctx->counters["sort.last_passes"] = npass;
        ctx->counters["sort.last_truncated"] = trunc.on ? 1 : 0;
    } else {
    // ---- 3. all digit histograms up front (one read per key column), then their exclusive scans ----
        rc = ctx->dalloc((void **)&d_hist, sizeof(unsigned long long) * 256 * npass);
        if (rc != HARK_OK) return cleanup_fail(rc, "");
        e = cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * 256 * npass, ctx->stream);
        for (int p0 = 0; p0 < npass && e == cudaSuccess;) {
            int p1 = p0;
            while (p1 < npass && passes[p1].key == passes[p0].key && p1 - p0 < MAXPASS) p1++;
            HistParams H;
            memset(&H, 0, sizeof H);
            const hk_sort_array &ka = arrays[keys[passes[p0].key].array];
            H.key = ka.in;
            H.n = n;
            H.npass = p1 - p0;
            for (int q = p0; q < p1; q++) H.f[q - p0] = passes[q].f;
            H.hist = d_hist + (size_t)256 * p0;
            const unsigned g = grid_for(ctx, n, 4);
            if (ka.width == 4) hk_hist_kernel<4><<<g, 256, 0, ctx->stream>>>(H);
            else hk_hist_kernel<8><<<g, 256, 0, ctx->stream>>>(H);
            e = cudaGetLastError();
            ctx->count_launch();
            p0 = p1;
        }
        if (e == cudaSuccess && hash_counts) { // bucket sizes for the caller (before the scan overwrites them)
            e = cudaMemcpyAsync(ctx->h_scalars, d_hist, sizeof(uint64_t) * 256, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e == cudaSuccess)
                for (int i = 0; i < hash_nparts; i++) hash_counts[i] = (int64_t)ctx->h_scalars[i];
        }
        if (e == cudaSuccess) {
            hk_hist_scan_kernel<<<npass, 256, 0, ctx->stream>>>(d_hist);
            e = cudaGetLastError();
            ctx->count_launch();
        }
        if (e != cudaSuccess) return cleanup_fail(HARK_ERR_CUDA, std::string("sort(hist): ") + cudaGetErrorString(e));
    
        // ---- 4. the passes ----
        const int64_t num_tiles = (n + STILE - 1) / STILE;
        uint64_t *d_status = nullptr; // [0] ticket per pass x npass ... then [num_tiles][256]
        const size_t status_words = (size_t)num_tiles * 256;
        rc = ctx->dalloc((void **)&d_status, sizeof(uint64_t) * (status_words + (size_t)npass));
        if (rc != HARK_OK) return cleanup_fail(rc, "");
        e = cudaMemsetAsync(d_status, 0, sizeof(uint64_t) * (status_words + (size_t)npass), ctx->stream);
        int occ4 = 0, occ8 = 0;
        const size_t smem = (size_t)STILE * 8;
        if (e == cudaSuccess) e = cudaFuncSetAttribute(hk_onesweep_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(hk_onesweep_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4, hk_onesweep_kernel<4>, ST, smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ8, hk_onesweep_kernel<8>, ST, smem);
        const int64_t want_occ = ctx->opt("sort.ctas_per_sm", 0);
    
        std::vector<const void *> cur(na);
        for (int a = 0; a < na; a++) cur[a] = arrays[a].in;
        for (int p = 0; p < npass && e == cudaSuccess; p++) {
            if (p > 0 && p % 255 == 0) // tags wrap: start over with a clean status array
                e = cudaMemsetAsync(d_status + npass, 0, sizeof(uint64_t) * status_words, ctx->stream);
            const int ob = p & 1;
            for (int a = 0; a < na && rc == HARK_OK; a++)
                if (!arrays[a].buf[ob]) rc = ctx->dalloc(&arrays[a].buf[ob], (size_t)n * arrays[a].width);
            if (rc != HARK_OK) {
                ctx->dfree(d_status);
                return cleanup_fail(rc, "");
            }
            PassParams P;
            memset(&P, 0, sizeof P);
            P.f = passes[p].f;
            P.na = na;
            P.ka = keys[passes[p].key].array;
            for (int a = 0; a < na; a++) {
                P.in[a] = cur[a];
                P.out[a] = arrays[a].buf[ob];
                P.width[a] = arrays[a].width;
            }
            P.n = n;
            P.num_tiles = num_tiles;
            P.status = d_status + npass;
            P.ticket = (unsigned long long *)(d_status + p);
            P.pass_offsets = d_hist + (size_t)256 * passes[p].hist_slot;
            P.tag = (uint32_t)(p % 255) + 1;
            const int kw = arrays[P.ka].width;
            int occ = std::max(1, kw == 4 ? occ4 : occ8);
            if (want_occ > 0) occ = (int)std::min<int64_t>(occ, want_occ);
            const unsigned grid = (unsigned)std::min<int64_t>(num_tiles, (int64_t)ctx->num_sms * occ);
            if (p == 0) ctx->kernel_begin();
            if (kw == 4) hk_onesweep_kernel<4><<<grid, ST, smem, ctx->stream>>>(P);
            else hk_onesweep_kernel<8><<<grid, ST, smem, ctx->stream>>>(P);
            e = cudaGetLastError();
            ctx->count_launch();
            if (p == npass - 1) ctx->kernel_end();
            for (int a = 0; a < na; a++) cur[a] = arrays[a].buf[ob];
        }
        ctx->dfree(d_status);
        if (e != cudaSuccess) return cleanup_fail(HARK_ERR_CUDA, std::string("sort(pass): ") + cudaGetErrorString(e));
    }
    const int fb = (npass - 1) & 1;
    for (auto &a : arrays) {
        a.result = a.buf[fb];
        a.buf[fb] = nullptr;
        if (a.buf[fb ^ 1]) {
            ctx->dfree(a.buf[fb ^ 1]);
            a.buf[fb ^ 1] = nullptr;
        }
    }
    ctx->dfree(d_hist);
    ctx->dfree(d_scratch);
    if (info) {
        int64_t wsum = 0;
        for (auto &a : arrays) wsum += a.width;
        info->bytes_moved = 2 * n * wsum * npass;
    }
    return HARK_OK;
}

// The K3t planning decision without a device (tests): bits[k] = significant bits of key k (most significant key first),
// sample = S normalised key tuples, row-major.  Returns 1 and fills kstar / q / shift / passes when the sort would run
// truncated, 0 when it keeps every pass.
extern "C" int hark_debug_plan_truncation(int64_t n, int32_t nk, const int32_t *bits, const uint64_t *sample, int32_t S,
                                          int32_t slack, int32_t *kstar, int32_t *q, int32_t *shift, int32_t *passes) {
    if (n < 2 || nk < 1 || nk > FIX_MAXK || !bits || (S > 0 && !sample)) return 0;
    int b[FIX_MAXK];
    for (int k = 0; k < nk; k++) b[k] = bits[k];
    TruncPlan tp;
    plan_from_sample(n, nk, b, sample, S, slack, &tp);
    if (!tp.on) return 0;
    int sh = 0;
    const int np = count_passes(b, tp.kstar, tp.q, &sh);
    if (kstar) *kstar = tp.kstar;
    if (q) *q = tp.q;
    if (shift) *shift = tp.shift;
    if (passes) *passes = np;
    return 1;
}
