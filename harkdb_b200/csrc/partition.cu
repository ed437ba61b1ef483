// partition.cu — K8a: one-pass range partition of (key, up to 3 four-byte value arrays) into <= 256 bins.
//
// Used by K2 (GROUP BY with more key slots than one shared-memory table holds: bin = (key - min) >> log2(K)) and
// by join + GROUP BY (bin = slice of the dimension lookup, so that the slice being probed stays L2-resident).
// The reference has no such step: it reaches the same end — equal keys adjacent — with 32 stable one-bit passes
// over whole rows (futhark/groupby.fut:8-22, join.fut:9-23).
//
// Three launches, no inter-CTA communication:
//   1. hk_part_hist_kernel   every CTA owns one contiguous CHUNK of tiles and counts its rows per bin
//                            (one streaming read of the key column, shared-memory atomics);
//   2. hk_part_scan_kernel   one CTA: per bin, exclusive scan over the chunks -> where each chunk's rows of each
//                            bin start; plus the 257 bin offsets the consumer needs;
//   3. hk_part_kernel        every CTA walks its chunk tile by tile: rank rows inside the tile with one
//                            shared-memory atomic each, reorder the tile in shared memory so every bin leaves as one
//                            contiguous run, advance the chunk's running per-bin offsets.  The next tile's loads
//                            are issued before the current tile is written out.  No look-back, no global atomics:
//                            the chunk's offsets were fixed by step 2.
// Order inside a bin is unspecified (callers aggregate).  HBM bytes: 4n (1.) + 2·n·(kw + 4·nv) (3.).
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

constexpr int PI = 8;            // rows per thread
constexpr int PG = PI / 4;       // 4-row groups per thread
constexpr int PMAXV = 3;
constexpr int PTILE_MIN = 2048;  // chunk boundaries are multiples of this (the smallest tile)

// digit of a row.  Integer keys only: the order key is raw ^ xmask (sign-bit flip for signed dtypes), and for
// 4-byte keys everything stays in 32-bit arithmetic (3 instructions).
struct DigitFn {
    uint64_t xmask; // order key = raw ^ xmask
    uint64_t base;  // smallest order key
    uint64_t last;  // span - 1 (rows with order key - base > last go to bin 0)
    int shift;
};

template <int KW>
__device__ __forceinline__ uint32_t part_digit(typename KRaw<KW>::T raw, const DigitFn &f) {
    if constexpr (KW == 4) {
        const uint32_t u = (raw ^ (uint32_t)f.xmask) - (uint32_t)f.base;
        return u <= (uint32_t)f.last ? (u >> f.shift) : 0u;
    } else {
        const uint64_t u = (raw ^ f.xmask) - f.base;
        return u <= f.last ? (uint32_t)(u >> f.shift) : 0u;
    }
}

struct PartParams {
    DigitFn f;
    const void *key_in;
    void *key_out;
    const uint32_t *val_in[PMAXV];
    uint32_t *val_out[PMAXV];
    int64_t n;
    int64_t num_tiles;
    int64_t tiles_per_chunk;
    int64_t tile_rows;                 // rows per tile of the scatter kernel (PT * PI)
    int num_chunks;
    uint32_t *chunk_counts;            // [num_chunks][256]
    unsigned long long *chunk_base;    // [num_chunks][256] global row where the chunk's rows of each bin start
    unsigned long long *offsets;       // [257] exclusive bin offsets
};

template <int KW>
__global__ void __launch_bounds__(1024) hk_part_hist_kernel(const __grid_constant__ PartParams P) {
    using T = typename KRaw<KW>::T;
    __shared__ uint32_t sh[256];
    if (threadIdx.x < 256) sh[threadIdx.x] = 0;
    __syncthreads();
    const T *p = reinterpret_cast<const T *>(P.key_in);
    constexpr int V = 16 / KW;
    const int64_t r0 = (int64_t)blockIdx.x * P.tiles_per_chunk * P.tile_rows;     // multiple of 2048
    const int64_t r1 = min(P.n, r0 + P.tiles_per_chunk * P.tile_rows);
    const int64_t nvec = (r1 - r0) / V;
    const T *q = p + r0;
#pragma unroll 4
    for (int64_t i = threadIdx.x; i < nvec; i += 1024) {
        T x[V];
        if constexpr (KW == 4) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(q) + i);
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        } else {
            const ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2 *>(q) + i);
            x[0] = v.x; x[1] = v.y;
        }
#pragma unroll
        for (int e = 0; e < V; e++) atomicAdd(&sh[part_digit<KW>(x[e], P.f)], 1u);
    }
    if (threadIdx.x < (int)((r1 - r0) - nvec * V)) atomicAdd(&sh[part_digit<KW>(q[nvec * V + threadIdx.x], P.f)], 1u);
    __syncthreads();
    if (threadIdx.x < 256) P.chunk_counts[(size_t)blockIdx.x * 256 + threadIdx.x] = sh[threadIdx.x];
}

// one CTA of 1024 threads: thread (q, b) scans quarter q of the chunks of bin b
__global__ void __launch_bounds__(1024) hk_part_scan_kernel(const __grid_constant__ PartParams P) {
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
    if (q == 0) { // exclusive scan of the 256 bin totals
        const int lane = b & 31, warp = b >> 5;
        unsigned long long inc = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(HK_FULL_MASK, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wtot[warp] = inc;
    }
    __syncthreads();
    if (q == 0) {
        const int lane = b & 31, warp = b >> 5;
        unsigned long long inc = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(HK_FULL_MASK, inc, o);
            if (lane >= o) inc += t;
        }
        unsigned long long off = 0;
        for (int w = 0; w < warp; w++) off += wtot[w];
        P.offsets[b] = off + inc - total;
        if (b == 255) P.offsets[256] = off + inc;
        s_q[0][b] = off + inc - total + 0ull; // bin start (quarter 0 starts here)
        // quarter starts: bin start + totals of the earlier quarters (s_q[1..3] still hold the quarter sums)
        const unsigned long long q1 = s_q[1][b], q2 = s_q[2][b], q0 = sum;
        const unsigned long long st = off + inc - total;
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

template <int KW, int NV, int PT>
__device__ __forceinline__ void part_load_tile(const PartParams &P, int64_t tile_base, int count, int tid,
                                               typename KRaw<KW>::T (&key)[PI], uint32_t (&val)[NV > 0 ? NV : 1][PI]) {
    using KT = typename KRaw<KW>::T;
    constexpr int PTILE = PT * PI;
    const KT *keyp = reinterpret_cast<const KT *>(P.key_in);
    if (count == PTILE) {
#pragma unroll
        for (int g = 0; g < PG; g++) {
            const int64_t r = tile_base + (int64_t)(g * PT + tid) * 4;
            if constexpr (KW == 4) {
                const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(keyp + r));
                key[g * 4 + 0] = v.x; key[g * 4 + 1] = v.y; key[g * 4 + 2] = v.z; key[g * 4 + 3] = v.w;
            } else {
                const ulonglong2 a = __ldcs(reinterpret_cast<const ulonglong2 *>(keyp + r));
                const ulonglong2 b = __ldcs(reinterpret_cast<const ulonglong2 *>(keyp + r + 2));
                key[g * 4 + 0] = a.x; key[g * 4 + 1] = a.y; key[g * 4 + 2] = b.x; key[g * 4 + 3] = b.y;
            }
        }
#pragma unroll
        for (int v = 0; v < NV; v++)
#pragma unroll
            for (int g = 0; g < PG; g++) {
                const int64_t r = tile_base + (int64_t)(g * PT + tid) * 4;
                const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(P.val_in[v] + r));
                val[v][g * 4 + 0] = x.x; val[v][g * 4 + 1] = x.y; val[v][g * 4 + 2] = x.z; val[v][g * 4 + 3] = x.w;
            }
    } else {
#pragma unroll
        for (int i = 0; i < PI; i++) {
            const int idx = ((i >> 2) * PT + tid) * 4 + (i & 3);
            key[i] = idx < count ? keyp[tile_base + idx] : (KT)0;
#pragma unroll
            for (int v = 0; v < NV; v++) val[v][i] = idx < count ? P.val_in[v][tile_base + idx] : 0u;
        }
    }
}

template <int KW, int NV, int PT>
__global__ void __launch_bounds__(PT, 1024 / PT) hk_part_kernel(const __grid_constant__ PartParams P) {
    using KT = typename KRaw<KW>::T;
    constexpr int PTILE = PT * PI;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    KT *s_key = reinterpret_cast<KT *>(s_dyn);
    uint32_t *s_val = reinterpret_cast<uint32_t *>(s_dyn + (size_t)PTILE * KW); // [NV][PTILE]
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_binstart[256];
    __shared__ uint64_t s_gbase[256];
    __shared__ uint64_t s_run[256];   // next global row of every bin for this chunk
    __shared__ uint32_t s_wtot[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * P.tiles_per_chunk;
    const int64_t t1 = min(P.num_tiles, t0 + P.tiles_per_chunk);
    if (tid < 256) {
        s_run[tid] = P.chunk_base[(size_t)blockIdx.x * 256 + tid];
        s_hist[tid] = 0;
    }
    if (t0 >= t1) return;

    KT key[PI];
    uint32_t val[NV > 0 ? NV : 1][PI];
    int count = (int)min((int64_t)PTILE, P.n - t0 * PTILE);
    part_load_tile<KW, NV, PT>(P, t0 * PTILE, count, tid, key, val);
    __syncthreads();

    for (int64_t tile = t0; tile < t1; tile++) {
        // ---- rank inside the tile: one shared-memory atomic per row (packed: digit << 16 | rank) ----
        uint32_t rd[PI];
#pragma unroll
        for (int i = 0; i < PI; i++) {
            const int idx = ((i >> 2) * PT + tid) * 4 + (i & 3);
            if (count == PTILE || idx < count) {
                const uint32_t d = part_digit<KW>(key[i], P.f);
                rd[i] = (d << 16) | atomicAdd(&s_hist[d], 1u);
            } else {
                rd[i] = 0xffffffffu;
            }
        }
        __syncthreads();
        // ---- thread b owns bin b: tile-local start, global base of the bin's run, advance the chunk offsets ----
        {
            const uint32_t sum = tid < 256 ? s_hist[tid] : 0u;
            if (tid < 256) s_hist[tid] = 0;
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
        // ---- tile-local reorder ----
#pragma unroll
        for (int i = 0; i < PI; i++) {
            if (rd[i] != 0xffffffffu) {
                const uint32_t pos = s_binstart[rd[i] >> 16] + (rd[i] & 0xffffu);
                s_key[pos] = key[i];
#pragma unroll
                for (int v = 0; v < NV; v++) s_val[v * PTILE + pos] = val[v][i];
            }
        }
        // ---- the registers are free: issue the next tile's loads before writing this one out ----
        const int cur_count = count;
        if (tile + 1 < t1) {
            count = (int)min((int64_t)PTILE, P.n - (tile + 1) * PTILE);
            part_load_tile<KW, NV, PT>(P, (tile + 1) * PTILE, count, tid, key, val);
        }
        __syncthreads();
        {
            KT *ko = reinterpret_cast<KT *>(P.key_out);
            if (cur_count == PTILE) {
#pragma unroll
                for (int i = 0; i < PI; i++) {
                    const int j = i * PT + tid;
                    const KT k = s_key[j];
                    const uint64_t g = s_gbase[part_digit<KW>(k, P.f)] + (uint64_t)j;
                    ko[g] = k;
#pragma unroll
                    for (int v = 0; v < NV; v++) P.val_out[v][g] = s_val[v * PTILE + j];
                }
            } else {
                for (int j = tid; j < cur_count; j += PT) {
                    const KT k = s_key[j];
                    const uint64_t g = s_gbase[part_digit<KW>(k, P.f)] + (uint64_t)j;
                    ko[g] = k;
#pragma unroll
                    for (int v = 0; v < NV; v++) P.val_out[v][g] = s_val[v * PTILE + j];
                }
            }
        }
        // the two barriers of the next iteration order these shared-memory reads before the next reorder
    }
}

template <int KW, int NV, int PT>
int part_occupancy(hark_ctx *ctx, int *occ) {
    const size_t smem = (size_t)PT * PI * (KW + 4 * NV);
    auto kern = hk_part_kernel<KW, NV, PT>;
    HK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, PT, smem));
    *occ = std::max(1, *occ);
    return HARK_OK;
}

template <int KW, int NV, int PT>
int launch_part(hark_ctx *ctx, const PartParams &P) {
    const size_t smem = (size_t)PT * PI * (KW + 4 * NV);
    hk_part_kernel<KW, NV, PT><<<(unsigned)P.num_chunks, PT, smem, ctx->stream>>>(P);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch();
    return HARK_OK;
}

// ------------------------------------------------------------------------------------------------
// K8t — tile-local partition with a directory (the default producer of K2's input since round 2).
//
// The consumers of a partition here only AGGREGATE, so rows of a bin need not be contiguous across the whole table —
// only findable.  Every 4096-row tile is therefore sorted by bin INSIDE shared memory and leaves as ONE contiguous,
// 16-byte aligned block of the scratch table (tile t -> rows [t·4096, (t+1)·4096)), rows stored array-of-structs
// ([key words | value words], RW = KW/4 + NV words per row), together with a directory word per (tile, bin) =
// start | end << 16 of the bin's run inside the tile.  Consequences against K8a above:
//   * no histogram pre-pass (one streaming read of the key column saved) and no scan kernel — the layout is static;
//   * the write-out is a TMA bulk copy shared -> global (cp.async.bulk, SASS UBLKCP.G.S) issued by one thread: no
//     per-row shared-memory read, no per-row address look-up, no per-row store instruction;
//   * one shared-memory store per row (the packed row) instead of one per carried array.
// HBM bytes: n·(KW + 4·NV) read + the same written + 4·nbins per tile of directory.  The consumer walks, for bin b,
// the segments (t, b) of all tiles (dense_agg.cu, hk_dagg_tiles_kernel).
// ------------------------------------------------------------------------------------------------
constexpr int TP_T = 512;             // threads per CTA, two CTAs per SM
constexpr int TP_TILE = TP_T * PI;    // rows per tile (4096: bin ends fit the directory's 16-bit halves)
static_assert(TP_TILE == HK_TPART_TILE, "tile size is part of the directory format");

struct TPartParams {
    DigitFn f;
    int hash;            // 1: bin = (hk_hash_key(raw) & hmask) >> f.shift  (slices of a hash table, join.cu)
    uint64_t hmask;
    const void *key_in;
    const uint32_t *val_in[PMAXV];
    uint32_t *rows_out;  // [num_tiles * TP_TILE][RW]
    uint32_t *dir;       // [num_tiles][nbins]
    int64_t n;
    int64_t num_tiles;
    int64_t tiles_per_cta;
    int nbins;
};

__device__ __forceinline__ void tp_bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)), "r"(s),
                 "r"(bytes)
                 : "memory");
}

// bin of a row with the partition constants in registers (HASH: slices of a hash table, else key ranges)
template <int KW, bool HASH>
__device__ __forceinline__ uint32_t tpart_digit(typename KRaw<KW>::T raw, typename KRaw<KW>::T xmask, typename KRaw<KW>::T base,
                                                typename KRaw<KW>::T last, int shift, uint64_t hmask) {
    if constexpr (HASH) {
        return (uint32_t)((hk_hash_key<KW>(raw) & hmask) >> shift);
    } else {
        const typename KRaw<KW>::T u = (raw ^ xmask) - base;
        return u <= last ? (uint32_t)(u >> shift) : 0u;
    }
}

// One tile: rank (one shared-memory atomic per row, all of a thread's atomics issued back to back), scan, scatter of the
// packed rows into the stage, request of the next tile, bulk copy out.  FULL = the tile has all TP_TILE rows (no bounds
// checks anywhere).
template <int KW, int NV, bool HASH, bool FULL>
__device__ __forceinline__ void tpart_tile(const TPartParams &P, const PartParams &L, uint32_t *s_stage, uint32_t *s_hist,
                                           uint32_t *s_binstart, uint32_t *s_wtot, uint32_t *s_hot, int64_t tile, int count, int next_count,
                                           typename KRaw<KW>::T (&key)[PI], uint32_t (&val)[NV > 0 ? NV : 1][PI],
                                           typename KRaw<KW>::T xmask, typename KRaw<KW>::T base, typename KRaw<KW>::T last, int shift) {
    using KT = typename KRaw<KW>::T;
    constexpr int RW = KW / 4 + NV;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t d[PI], rk[PI];
#pragma unroll
    for (int i = 0; i < PI; i++) d[i] = tpart_digit<KW, HASH>(key[i], xmask, base, last, shift, P.hmask);
    // Skewed keys: when one bin took more than a quarter of the previous tile, its rows are ranked with one vote and ONE
    // atomic per warp instruction instead of one same-address atomic per row (which the hardware serialises: with Zipf
    // keys 70 % of a tile lands in one bin).  `hot` is uniform over the CTA, so the branch does not diverge.
    const uint32_t hot = *s_hot;
    if (hot != 0xffffffffu) {
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int i = 0; i < PI; i++) {
            const int idx = ((i >> 2) * TP_T + tid) * 4 + (i & 3);
            const bool valid = FULL || idx < count;
            const bool is_hot = valid && d[i] == hot;
            const uint32_t m = __ballot_sync(HK_FULL_MASK, is_hot);
            uint32_t base = 0;
            if (m) {
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(&s_hist[hot], (uint32_t)__popc(m));
                base = __shfl_sync(HK_FULL_MASK, base, leader);
            }
            if (is_hot) rk[i] = base + __popc(m & lt);
            else if (valid) rk[i] = atomicAdd(&s_hist[d[i]], 1u);
        }
    } else {
#pragma unroll
        for (int i = 0; i < PI; i++) {
            const int idx = ((i >> 2) * TP_T + tid) * 4 + (i & 3);
            if (FULL || idx < count) rk[i] = atomicAdd(&s_hist[d[i]], 1u);
        }
    }
    // the bulk copy of the previous tile must have finished READING the stage before anyone overwrites it
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    // ---- thread b owns bin b: start of the bin's run inside the tile, directory word ----
    {
        const uint32_t sum = tid < 256 ? s_hist[tid] : 0u;
        if (tid < 256) s_hist[tid] = 0;
        const uint32_t inc = hk_warp_incl_scan_u32(sum);
        if (lane == 31 && warp < 8) s_wtot[warp] = inc;
        __syncthreads();
        if (tid < 256) {
            uint32_t woff = 0;
            for (int w = 0; w < warp; w++) woff += s_wtot[w];
            const uint32_t binstart = woff + inc - sum;
            s_binstart[tid] = binstart;
            if (tid < P.nbins) P.dir[(size_t)tile * P.nbins + tid] = binstart | ((binstart + sum) << 16);
            // the fullest bin of this tile (count << 8 | bin), for the next tile's ranking
            uint32_t best = (sum << 8) | (uint32_t)tid;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(HK_FULL_MASK, best, o));
            if (lane == 0) s_wtot[8 + warp] = best;
        }
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t best = 0;
        for (int w = 0; w < 8; w++) best = max(best, s_wtot[8 + w]);
        *s_hot = (best >> 8) > (uint32_t)(TP_TILE / 4) ? (best & 0xffu) : 0xffffffffu; // read after the next barriers
    }
    // ---- the packed rows go to their slot of the stage ----
#pragma unroll
    for (int i = 0; i < PI; i++) {
        const int idx = ((i >> 2) * TP_T + tid) * 4 + (i & 3);
        if (FULL || idx < count) {
            const uint32_t pos = s_binstart[d[i]] + rk[i];
            uint32_t w[RW];
            if constexpr (KW == 4) {
                w[0] = key[i];
            } else {
                w[0] = (uint32_t)key[i];
                w[1] = (uint32_t)(key[i] >> 32);
            }
#pragma unroll
            for (int v = 0; v < NV; v++) w[KW / 4 + v] = val[v][i];
            if constexpr (RW == 2) {
                reinterpret_cast<uint2 *>(s_stage)[pos] = make_uint2(w[0], w[1]);
            } else if constexpr (RW == 4) {
                reinterpret_cast<uint4 *>(s_stage)[pos] = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
#pragma unroll
                for (int q = 0; q < RW; q++) s_stage[pos * RW + q] = w[q];
            }
        }
    }
    // ---- the registers are free: request the next tile before this one leaves ----
    if (next_count > 0) part_load_tile<KW, NV, TP_T>(L, (tile + 1) * TP_TILE, next_count, tid, key, val);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy stores -> visible to the bulk copy
    __syncthreads();
    if (tid == 0) {
        // the whole tile block is written (rows past a ragged last tile's count are never read: the directory bounds
        // every run); 16 KB pieces
        constexpr uint32_t BYTES = (uint32_t)TP_TILE * RW * 4;
        unsigned char *g = reinterpret_cast<unsigned char *>(P.rows_out) + (size_t)tile * BYTES;
        const unsigned char *sm = reinterpret_cast<const unsigned char *>(s_stage);
#pragma unroll 1
        for (uint32_t o = 0; o < BYTES; o += 16384u) tp_bulk_store(g + o, sm + o, min(16384u, BYTES - o));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}

template <int KW, int NV, bool HASH>
__global__ void __launch_bounds__(TP_T, 3) hk_tpart_kernel(const __grid_constant__ TPartParams P) {
    using KT = typename KRaw<KW>::T;
    constexpr int RW = KW / 4 + NV;
    extern __shared__ __align__(128) uint32_t s_dyn32[]; // stage (TP_TILE rows x RW words), then the small arrays
    uint32_t *s_stage = s_dyn32;
    uint32_t *s_hist = s_dyn32 + (size_t)TP_TILE * RW;
    uint32_t *s_binstart = s_hist + 256;
    uint32_t *s_wtot = s_binstart + 256; // [0..8) warp totals of the scan, [8..16) warp maxima
    uint32_t *s_hot = s_wtot + 16;       // bin that held more than a quarter of the previous tile, or 0xffffffff

    const int tid = threadIdx.x;
    if (tid == 0) *s_hot = 0xffffffffu;
    const int64_t t0 = (int64_t)blockIdx.x * P.tiles_per_cta;
    const int64_t t1 = min(P.num_tiles, t0 + P.tiles_per_cta);
    if (tid < 256) s_hist[tid] = 0;
    if (t0 >= t1) return;

    PartParams L; // the loader of K8a, reused: it only looks at the input pointers
    L.key_in = P.key_in;
    for (int v = 0; v < PMAXV; v++) L.val_in[v] = P.val_in[v];
    const KT xmask = (KT)P.f.xmask, base = (KT)P.f.base, last = (KT)P.f.last;
    const int shift = P.f.shift;
    KT key[PI];
    uint32_t val[NV > 0 ? NV : 1][PI];
    int count = (int)min((int64_t)TP_TILE, P.n - t0 * TP_TILE);
    part_load_tile<KW, NV, TP_T>(L, t0 * TP_TILE, count, tid, key, val);
    __syncthreads();

    for (int64_t tile = t0; tile < t1; tile++) {
        const int next_count = tile + 1 < t1 ? (int)min((int64_t)TP_TILE, P.n - (tile + 1) * TP_TILE) : 0;
        if (count == TP_TILE)
            tpart_tile<KW, NV, HASH, true>(P, L, s_stage, s_hist, s_binstart, s_wtot, s_hot, tile, count, next_count, key, val, xmask, base, last, shift);
        else
            tpart_tile<KW, NV, HASH, false>(P, L, s_stage, s_hist, s_binstart, s_wtot, s_hot, tile, count, next_count, key, val, xmask, base, last, shift);
        count = next_count;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // shared memory must outlive the copy
}

template <int KW, int NV>
int launch_tpart(hark_ctx *ctx, const TPartParams &P, unsigned grid) {
    const size_t smem = (size_t)TP_TILE * (KW / 4 + NV) * 4 + (256 + 256 + 16 + 4) * 4;
    void (*kern)(const TPartParams) = P.hash ? hk_tpart_kernel<KW, NV, true> : hk_tpart_kernel<KW, NV, false>;
    HK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, TP_T, smem, ctx->stream>>>(P);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch();
    return HARK_OK;
}

#define HK_PART_DISPATCH_NV(fn, KWv, PTv, ...)                                                                  \
    (nv == 0 ? fn<KWv, 0, PTv>(__VA_ARGS__) : nv == 1 ? fn<KWv, 1, PTv>(__VA_ARGS__)                            \
                                            : nv == 2 ? fn<KWv, 2, PTv>(__VA_ARGS__) : fn<KWv, 3, PTv>(__VA_ARGS__))
#define HK_PART_DISPATCH(fn, ...)                                                                               \
    (pt == 512 ? (kw == 4 ? HK_PART_DISPATCH_NV(fn, 4, 512, __VA_ARGS__) : HK_PART_DISPATCH_NV(fn, 8, 512, __VA_ARGS__)) \
               : (kw == 4 ? HK_PART_DISPATCH_NV(fn, 4, 256, __VA_ARGS__) : HK_PART_DISPATCH_NV(fn, 8, 256, __VA_ARGS__)))

} // namespace

int hk_partition_pass(hark_ctx *ctx, int64_t n, const void *key, int kw, const hk_part_spec &spec, int nv,
                      const void *const *vals, void **key_out, void **vals_out, unsigned long long **d_offsets) {
    if (nv > PMAXV || nv < 0 || spec.nbins < 1 || spec.nbins > 256 || (kw != 4 && kw != 8))
        return ctx->fail(HARK_ERR_UNSUPPORTED, "partition_pass: unsupported shape");
    *key_out = nullptr;
    *d_offsets = nullptr;
    for (int v = 0; v < nv; v++) vals_out[v] = nullptr;
    struct Owner { // frees everything unless released
        hark_ctx *ctx;
        std::vector<void *> v;
        ~Owner() {
            for (void *p : v) ctx->dfree(p);
        }
    } own{ctx, {}}, tmp{ctx, {}};
    auto alloc_into = [&](Owner &o, void **p, size_t bytes) {
        int rc = ctx->dalloc(p, bytes);
        if (rc == HARK_OK) o.v.push_back(*p);
        return rc;
    };

    if (!hk_dtype_int(spec.dtype)) return ctx->fail(HARK_ERR_UNSUPPORTED, "partition_pass: integer keys only");
    PartParams P;
    memset(&P, 0, sizeof P);
    P.f.xmask = spec.dtype == HARK_I32 ? 0x80000000ull : spec.dtype == HARK_I64 ? 0x8000000000000000ull : 0ull;
    P.f.base = spec.base;
    P.f.last = spec.span == 0 ? ~0ull : spec.span - 1;
    if (kw == 4 && P.f.last > 0xffffffffull) P.f.last = 0xffffffffull;
    P.f.shift = spec.shift;
    P.key_in = key;
    P.n = n;
    const int pt = ctx->opt("part.threads", 512) == 256 ? 256 : 512; // CTA size: 2048- or 4096-row tiles
    P.tile_rows = (int64_t)pt * PI;
    P.num_tiles = (n + P.tile_rows - 1) / P.tile_rows;
    int occ = 1;
    HK_TRY(HK_PART_DISPATCH(part_occupancy, ctx, &occ));
    const int64_t want = ctx->opt("part.ctas_per_sm", 0);
    if (want > 0) occ = (int)std::min<int64_t>(occ, want);
    const int64_t max_chunks = (int64_t)ctx->num_sms * occ;
    P.tiles_per_chunk = std::max<int64_t>(1, (P.num_tiles + max_chunks - 1) / max_chunks);
    P.num_chunks = (int)std::max<int64_t>(1, (P.num_tiles + P.tiles_per_chunk - 1) / P.tiles_per_chunk);

    HK_TRY(alloc_into(own, (void **)&P.offsets, 257 * sizeof(unsigned long long)));
    HK_TRY(alloc_into(tmp, (void **)&P.chunk_counts, (size_t)P.num_chunks * 256 * sizeof(uint32_t)));
    HK_TRY(alloc_into(tmp, (void **)&P.chunk_base, (size_t)P.num_chunks * 256 * sizeof(unsigned long long)));
    HK_TRY(alloc_into(own, &P.key_out, (size_t)std::max<int64_t>(n, 1) * kw));
    for (int v = 0; v < nv; v++) {
        void *vo = nullptr;
        HK_TRY(alloc_into(own, &vo, (size_t)std::max<int64_t>(n, 1) * 4));
        P.val_in[v] = (const uint32_t *)vals[v];
        P.val_out[v] = (uint32_t *)vo;
    }
    if (kw == 4) hk_part_hist_kernel<4><<<(unsigned)P.num_chunks, 1024, 0, ctx->stream>>>(P);
    else hk_part_hist_kernel<8><<<(unsigned)P.num_chunks, 1024, 0, ctx->stream>>>(P);
    HK_CHECK_LAUNCH(ctx);
    hk_part_scan_kernel<<<1, 1024, 0, ctx->stream>>>(P);
    HK_CHECK_LAUNCH(ctx);
    ctx->count_launch(2);
    if (n > 0) HK_TRY(HK_PART_DISPATCH(launch_part, ctx, P));
    *key_out = P.key_out;
    for (int v = 0; v < nv; v++) vals_out[v] = P.val_out[v];
    *d_offsets = P.offsets;
    own.v.clear(); // ownership moves to the caller
    return HARK_OK;
}

int hk_tile_partition(hark_ctx *ctx, int64_t n, const void *key, int kw, const hk_part_spec &spec, uint64_t hash_mask, int nv,
                      const void *const *vals, hk_tpart *out) {
    if (nv > PMAXV || nv < 0 || spec.nbins < 1 || spec.nbins > 256 || (kw != 4 && kw != 8) || n <= 0)
        return ctx->fail(HARK_ERR_UNSUPPORTED, "tile_partition: unsupported shape");
    if (!hk_dtype_int(spec.dtype)) return ctx->fail(HARK_ERR_UNSUPPORTED, "tile_partition: integer keys only");
    *out = hk_tpart{};
    TPartParams P;
    memset(&P, 0, sizeof P);
    P.f.xmask = spec.dtype == HARK_I32 ? 0x80000000ull : spec.dtype == HARK_I64 ? 0x8000000000000000ull : 0ull;
    P.f.base = spec.base;
    P.f.last = spec.span == 0 ? ~0ull : spec.span - 1;
    if (kw == 4 && P.f.last > 0xffffffffull) P.f.last = 0xffffffffull;
    P.f.shift = spec.shift;
    P.hash = hash_mask != 0;
    P.hmask = hash_mask;
    P.key_in = key;
    for (int v = 0; v < nv; v++) P.val_in[v] = (const uint32_t *)vals[v];
    P.n = n;
    P.nbins = spec.nbins;
    P.num_tiles = (n + TP_TILE - 1) / TP_TILE;
    const int rw = kw / 4 + nv;
    const int64_t ctas_per_sm = std::min<int64_t>(ctx->opt("tpart.ctas_per_sm", 3), rw <= 4 ? 3 : 2);
    const int64_t max_ctas = (int64_t)ctx->num_sms * std::max<int64_t>(1, ctas_per_sm);
    P.tiles_per_cta = std::max<int64_t>(1, (P.num_tiles + max_ctas - 1) / max_ctas);
    const unsigned grid = (unsigned)((P.num_tiles + P.tiles_per_cta - 1) / P.tiles_per_cta);
    void *rows = nullptr, *dir = nullptr;
    HK_TRY(ctx->dalloc(&rows, (size_t)P.num_tiles * TP_TILE * rw * 4));
    int rc = ctx->dalloc(&dir, (size_t)P.num_tiles * spec.nbins * 4);
    if (rc != HARK_OK) {
        ctx->dfree(rows);
        return rc;
    }
    P.rows_out = (uint32_t *)rows;
    P.dir = (uint32_t *)dir;
#define HK_TPART_NV(KWv) (nv == 0 ? launch_tpart<KWv, 0>(ctx, P, grid) : nv == 1 ? launch_tpart<KWv, 1>(ctx, P, grid) \
                          : nv == 2 ? launch_tpart<KWv, 2>(ctx, P, grid) : launch_tpart<KWv, 3>(ctx, P, grid))
    rc = kw == 4 ? HK_TPART_NV(4) : HK_TPART_NV(8);
#undef HK_TPART_NV
    if (rc != HARK_OK) {
        ctx->dfree(rows);
        ctx->dfree(dir);
        return rc;
    }
    out->rows = (uint32_t *)rows;
    out->dir = (uint32_t *)dir;
    out->num_tiles = P.num_tiles;
    out->nbins = spec.nbins;
    out->rw = rw;
    return HARK_OK;
}
