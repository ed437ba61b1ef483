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

// ------------------------------------------------------------------------------------------------
// hash build (sparse keys): open addressing with linear probing over hk_hash_key(key) & hmask, load factor <= 1/2.
//   4-byte keys: 8-byte entries {key, payload + 1}; an entry is claimed with ONE 64-bit compare-and-swap, so a second
//                row with the same key meets the first one's entry and is reported as a duplicate on the spot;
//   8-byte keys: 16-byte entries {key, payload + 1 | (row + 1) << 32}; the payload word is claimed by CAS, the key is
//                stored after it, and hk_hash_verify_kernel (a second launch) checks that every row finds ITSELF
//                first — a duplicate key finds the other row.
// payload = group slot (ordkey(g) - g_lo) for the fused probe + aggregate, or the dimension row for the
// materialising path.  The probe side is hash_probe in dense_agg.cu / hk_hash_probe_kernel below.
// ------------------------------------------------------------------------------------------------
template <int KW>
__device__ __forceinline__ uint32_t build_payload(const void *g, int g_dtype, unsigned long long g_lo, int64_t i, int payload_row) {
    if (payload_row) return (uint32_t)i;
    unsigned long long ord;
    if (g_dtype == HARK_I64) ord = hk_ordkey64(reinterpret_cast<const unsigned long long *>(g)[i], g_dtype);
    else ord = hk_ordkey32(reinterpret_cast<const uint32_t *>(g)[i], g_dtype);
    return (uint32_t)(ord - g_lo);
}

template <int KW>
__global__ void __launch_bounds__(256) hk_hash_build_kernel(const void *pk, const void *g, int g_dtype, int64_t n_dim,
                                                             unsigned long long g_lo, int payload_row, void *htab,
                                                             unsigned long long hmask, unsigned int *dup_flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_dim; i += stride) {
        const uint32_t pay1 = build_payload<KW>(g, g_dtype, g_lo, i, payload_row) + 1u;
        if constexpr (KW == 4) {
            const uint32_t key = reinterpret_cast<const uint32_t *>(pk)[i];
            unsigned long long *t = reinterpret_cast<unsigned long long *>(htab);
            const unsigned long long packed = (unsigned long long)key | ((unsigned long long)pay1 << 32);
            unsigned long long h = hk_hash_key<4>(key) & hmask;
            while (true) {
                const unsigned long long old = atomicCAS(t + h, 0ull, packed);
                if (old == 0ull) break;
                if ((uint32_t)old == key) {
                    *dup_flag = 1u;
                    break;
                }
                h = (h + 1) & hmask;
            }
        } else {
            const unsigned long long key = reinterpret_cast<const unsigned long long *>(pk)[i];
            unsigned long long *t = reinterpret_cast<unsigned long long *>(htab); // entry h = words 2h (key), 2h+1 (payload)
            const unsigned long long packed = (unsigned long long)pay1 | ((unsigned long long)(i + 1) << 32);
            unsigned long long h = hk_hash_key<8>(key) & hmask;
            while (true) {
                const unsigned long long old = atomicCAS(t + 2 * h + 1, 0ull, packed);
                if (old == 0ull) {
                    t[2 * h] = key;
                    break;
                }
                h = (h + 1) & hmask;
            }
        }
    }
}

// The same build over K8t's output (dimension rows [pk, g] sorted by hash-table SLICE inside every tile): all warps walk
// slice 0 of every tile, then slice 1, ... so the compare-and-swaps of a slice land while it is L2-resident instead of
// being 10^8 random atomics over a table of gigabytes.  4-byte keys, 4-byte group column, group-slot payload.
__global__ void __launch_bounds__(256) hk_hash_build_tiles_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ dir,
                                                                   long long num_tiles, int nbins, int g_dtype, unsigned long long g_lo,
                                                                   void *htab, unsigned long long hmask, unsigned int *dup_flag,
                                                                   unsigned long long *ticket) {
    unsigned long long *t = reinterpret_cast<unsigned long long *>(htab);
    const int lane = threadIdx.x & 31;
    // (slice, tile) units in slice-major order from a ticket counter, 16 at a time: every warp of the chip works on the
    // same one or two slices (a static assignment lets the CTAs drift apart over the 128 slices)
    const long long total = (long long)nbins * num_tiles;
    while (true) {
        long long u0 = 0;
        if (lane == 0) u0 = (long long)atomicAdd(ticket, 16ull);
        u0 = __shfl_sync(HK_FULL_MASK, u0, 0);
        if (u0 >= total) break;
        const long long u1 = min(total, u0 + 16);
        for (long long uu = u0; uu < u1; uu++) {
            const int b = (int)(uu / num_tiles);
            const long long u = uu - (long long)b * num_tiles;
            const uint32_t w = __ldg(dir + (size_t)u * nbins + b);
            const uint32_t s = w & 0xffffu, e = w >> 16;
            for (uint32_t r = s + lane; r < e; r += 32) {
                const uint2 row = __ldcs(reinterpret_cast<const uint2 *>(rows) + (size_t)u * HK_TPART_TILE + r);
                const uint32_t key = row.x;
                const uint32_t pay1 = (uint32_t)((unsigned long long)hk_ordkey32(row.y, g_dtype) - g_lo) + 1u;
                const unsigned long long packed = (unsigned long long)key | ((unsigned long long)pay1 << 32);
                unsigned long long h = hk_hash_key<4>(key) & hmask;
                while (true) {
                    const unsigned long long old = atomicCAS(t + h, 0ull, packed);
                    if (old == 0ull) break;
                    if ((uint32_t)old == key) {
                        *dup_flag = 1u;
                        break;
                    }
                    h = (h + 1) & hmask;
                }
            }
        }
    }
}

// 8-byte keys: every row must be the first entry with its key on its probe sequence
__global__ void __launch_bounds__(256) hk_hash_verify_kernel(const void *pk, int64_t n_dim, const void *htab, unsigned long long hmask,
                                                              unsigned int *dup_flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const ulonglong2 *t = reinterpret_cast<const ulonglong2 *>(htab);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_dim; i += stride) {
        const unsigned long long key = reinterpret_cast<const unsigned long long *>(pk)[i];
        unsigned long long h = hk_hash_key<8>(key) & hmask;
        while (true) {
            const ulonglong2 e = t[h];
            if ((uint32_t)e.y == 0u) { // cannot happen for an inserted key; treat as corruption
                *dup_flag = 1u;
                break;
            }
            if (e.x == key) {
                if ((e.y >> 32) != (unsigned long long)(i + 1)) *dup_flag = 1u;
                break;
            }
            h = (h + 1) & hmask;
        }
    }
}

// materialising probe: gkey[i] = dim.g[row(fk[i])], hit[i] = 1 when fk[i] has a match
template <int KW>
__global__ void __launch_bounds__(256) hk_hash_probe_kernel(const void *fk, int64_t n_fact, const void *htab, unsigned long long hmask,
                                                             const void *g, int g_width, void *gkey_out, uint32_t *hit_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_fact; i += stride) {
        uint32_t row = 0xffffffffu;
        if constexpr (KW == 4) {
            const uint32_t key = reinterpret_cast<const uint32_t *>(fk)[i];
            const uint2 *t = reinterpret_cast<const uint2 *>(htab);
            unsigned long long h = hk_hash_key<4>(key) & hmask;
            while (true) {
                const uint2 e = __ldg(t + h);
                if (e.y == 0u) break;
                if (e.x == key) {
                    row = e.y - 1u;
                    break;
                }
                h = (h + 1) & hmask;
            }
        } else {
            const unsigned long long key = reinterpret_cast<const unsigned long long *>(fk)[i];
            const ulonglong2 *t = reinterpret_cast<const ulonglong2 *>(htab);
            unsigned long long h = hk_hash_key<8>(key) & hmask;
            while (true) {
                const ulonglong2 e = __ldg(t + h);
                if ((uint32_t)e.y == 0u) break;
                if (e.x == key) {
                    row = (uint32_t)e.y - 1u;
                    break;
                }
                h = (h + 1) & hmask;
            }
        }
        const bool hit = row != 0xffffffffu;
        hit_out[i] = hit;
        if (g_width == 4) reinterpret_cast<uint32_t *>(gkey_out)[i] = hit ? reinterpret_cast<const uint32_t *>(g)[row] : 0u;
        else reinterpret_cast<uint64_t *>(gkey_out)[i] = hit ? reinterpret_cast<const uint64_t *>(g)[row] : 0ull;
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

} // namespace

// ================================================================================================
// Plain inner equi-join, typed (hk_join_ex).  Two strategies:
//   order = 1  reference order (join.fut:55-75: key ascending, then left row, then right row): both sides are sorted
//              as (key, row id) by K3; hk_mj_bounds_kernel walks tiles of the sorted left side, narrows the right
//              side to the tile's key range with two searches per CTA, finds every left row's match range inside
//              it, and turns the counts into output offsets with ONE decoupled look-back per tile; the expand kernel
//              is balanced over OUTPUT rows (a CTA owns a fixed slice of the result, whatever the duplication).
//   order = 0  hash build on db2 + probe with db1 in row order (multiset result, matches of one left row in
//              unspecified order): count -> look-back scan per tile -> expand; nothing is sorted, the probe side
//              needs no row ids and may exceed 2^32 rows.
// ================================================================================================
namespace {

constexpr int MJ_T = 256, MJ_I = 8, MJ_TILE = MJ_T * MJ_I;

template <int KW> struct JRaw;
template <> struct JRaw<4> { using T = uint32_t; };
template <> struct JRaw<8> { using T = uint64_t; };

template <int KW>
__device__ __forceinline__ uint64_t j_ord(typename JRaw<KW>::T raw, int dtype) {
    if constexpr (KW == 4) return (uint64_t)hk_ordkey32(raw, dtype);
    else return hk_ordkey64(raw, dtype);
}

// block-wide exclusive scan of one u64 per thread (MJ_T threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ unsigned long long mj_block_excl_scan(unsigned long long v, unsigned long long *s_w /* [MJ_T/32] */,
                                                                 unsigned long long *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(HK_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    unsigned long long off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < MJ_T / 32; w++) {
        const unsigned long long x = s_w[w];
        if (w < warp) off += x;
        tot += x;
    }
    *total = tot;
    __syncthreads();
    return off + inc - v;
}

struct MjParams {
    const void *k1, *k2; // sorted keys
    int64_t n1, n2;
    int dtype;           // order of the keys (HARK_U32 for the pinned entry)
    uint32_t *lb;        // [n1] first matching position in k2
    unsigned long long *offs; // [n1 + 1] exclusive output offsets
    uint64_t *state;     // look-back words, one per tile
    unsigned long long *ticket;
    int64_t num_tiles;
    const long long *ranges; // [num_tiles][2] slice of k2 that can match the tile (hk_mj_tile_ranges_kernel)
};

// The slice of the sorted right side that can match left tile t: [lower_bound(first key), upper_bound(last key)).  One
// thread per tile boundary: the ~24 dependent loads of a search over n2 run for all tiles at once instead of inside
// every tile with 254 threads waiting at a barrier (ncu, round 2: barrier = 11.3 warps per issue in the bounds kernel).
template <int KW>
__global__ void __launch_bounds__(256) hk_mj_tile_ranges_kernel(const void *k1v, const void *k2v, int64_t n1, int64_t n2, int dtype,
                                                                int64_t num_tiles, long long *ranges /* [num_tiles][2] */) {
    using KT = typename JRaw<KW>::T;
    const KT *k1 = reinterpret_cast<const KT *>(k1v), *k2 = reinterpret_cast<const KT *>(k2v);
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= 2 * num_tiles) return;
    const int64_t tile = q >> 1;
    const bool upper = q & 1;
    const int64_t i = upper ? min(n1, (tile + 1) * (int64_t)MJ_TILE) - 1 : tile * (int64_t)MJ_TILE;
    const uint64_t key = j_ord<KW>(k1[i], dtype);
    int64_t lo = 0, hi = n2;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const uint64_t x = j_ord<KW>(k2[mid], dtype);
        if (upper ? (x <= key) : (x < key)) lo = mid + 1; else hi = mid;
    }
    ranges[q] = lo;
}

constexpr int MJ_STAGE = 4096; // right-side keys of a tile's slice staged in shared memory (order keys, 8 bytes each)

template <int KW>
__global__ void __launch_bounds__(MJ_T) hk_mj_bounds_kernel(const __grid_constant__ MjParams P) {
    using KT = typename JRaw<KW>::T;
    __shared__ unsigned long long s_w[MJ_T / 32];
    __shared__ long long s_tile;
    __shared__ unsigned long long s_excl;
    __shared__ uint64_t s_k2[MJ_STAGE];
    const KT *k1 = reinterpret_cast<const KT *>(P.k1), *k2 = reinterpret_cast<const KT *>(P.k2);
    while (true) {
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(P.ticket, 1ull);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.num_tiles) break;
        const int64_t i0 = tile * MJ_TILE, i1 = min(P.n1, i0 + MJ_TILE);
        const int64_t L = P.ranges[2 * tile], U = P.ranges[2 * tile + 1];
        const bool staged = U - L <= MJ_STAGE; // the usual case: the searches of the tile's rows run in shared memory
        if (staged)
            for (int64_t j = threadIdx.x; j < U - L; j += MJ_T) s_k2[j] = j_ord<KW>(k2[L + j], P.dtype);
        __syncthreads();
        uint32_t lbv[MJ_I];
        unsigned long long cnt[MJ_I], mine = 0;
#pragma unroll
        for (int e = 0; e < MJ_I; e++) {
            const int64_t i = i0 + (int64_t)threadIdx.x * MJ_I + e; // thread-contiguous rows: offsets come out in row order
            lbv[e] = 0;
            cnt[e] = 0;
            if (i < i1) {
                const uint64_t key = j_ord<KW>(k1[i], P.dtype);
                int64_t lb, ub;
                if (staged) {
                    int lo = 0, hi = (int)(U - L);
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (s_k2[mid] < key) lo = mid + 1; else hi = mid;
                    }
                    lb = L + lo;
                    hi = (int)(U - L);
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (s_k2[mid] <= key) lo = mid + 1; else hi = mid;
                    }
                    ub = L + lo;
                } else {
                    int64_t lo = L, hi = U;
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        if (j_ord<KW>(k2[mid], P.dtype) < key) lo = mid + 1; else hi = mid;
                    }
                    lb = lo;
                    hi = U;
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        if (j_ord<KW>(k2[mid], P.dtype) <= key) lo = mid + 1; else hi = mid;
                    }
                    ub = lo;
                }
                lbv[e] = (uint32_t)lb;
                cnt[e] = (unsigned long long)(ub - lb);
            }
            mine += cnt[e];
        }
        unsigned long long total;
        const unsigned long long excl_in_tile = mj_block_excl_scan(mine, s_w, &total);
        if (threadIdx.x < 32) {
            const unsigned long long ex = hk_lookback_u64(P.state, tile, total);
            if (threadIdx.x == 0) s_excl = ex;
        }
        __syncthreads();
        unsigned long long run = s_excl + excl_in_tile;
#pragma unroll
        for (int e = 0; e < MJ_I; e++) {
            const int64_t i = i0 + (int64_t)threadIdx.x * MJ_I + e;
            if (i < i1) {
                P.lb[i] = lbv[e];
                P.offs[i] = run;
                run += cnt[e];
            }
        }
        if (tile == P.num_tiles - 1 && threadIdx.x == MJ_T - 1) P.offs[P.n1] = s_excl + total;
        __syncthreads();
    }
}

struct JoinCols {
    int l, k;
    const void *src1[JMAXC];
    const void *src2[JMAXC];
    int w1[JMAXC], w2[JMAXC];
    void *dst[2 * JMAXC];
};

__device__ __forceinline__ void join_emit(const JoinCols &C, int64_t p, int64_t r1, int64_t r2) {
    for (int c = 0; c < C.l; c++) {
        if (C.w1[c] == 4) reinterpret_cast<uint32_t *>(C.dst[c])[p] = reinterpret_cast<const uint32_t *>(C.src1[c])[r1];
        else reinterpret_cast<uint64_t *>(C.dst[c])[p] = reinterpret_cast<const uint64_t *>(C.src1[c])[r1];
    }
    for (int c = 0; c < C.k; c++) {
        if (C.w2[c] == 4) reinterpret_cast<uint32_t *>(C.dst[C.l + c])[p] = reinterpret_cast<const uint32_t *>(C.src2[c])[r2];
        else reinterpret_cast<uint64_t *>(C.dst[C.l + c])[p] = reinterpret_cast<const uint64_t *>(C.src2[c])[r2];
    }
}

struct MjExpandParams {
    int64_t P, n1;
    const unsigned long long *offs; // [n1 + 1]
    const uint32_t *lb;
    const uint32_t *rid1, *rid2;    // rid1 == null: the left columns in C were carried through the sort (index = sorted position)
    int64_t slice;                  // output rows per CTA iteration
    JoinCols C;
};

// balanced over the OUTPUT: a CTA owns result rows [q * slice, (q + 1) * slice), finds the left rows that produce them
// with two searches, and every thread then searches only inside that (small, cached) range
constexpr int MJ_CACHE = 6144; // left rows of a slice whose offsets are staged in shared memory

__global__ void __launch_bounds__(256) hk_mj_expand_kernel(const __grid_constant__ MjExpandParams E) {
    __shared__ long long s_r[2];
    __shared__ uint32_t s_off[MJ_CACHE]; // offs[ia + k] - p0, saturated (rows past the slice compare as "greater")
    const int64_t nslices = (E.P + E.slice - 1) / E.slice;
    for (int64_t q = blockIdx.x; q < nslices; q += gridDim.x) {
        const int64_t p0 = q * E.slice, p1 = min(E.P, p0 + E.slice);
        if (threadIdx.x < 2) { // last i with offs[i] <= p, for p = p0 and p = p1 - 1
            const unsigned long long p = (unsigned long long)(threadIdx.x == 0 ? p0 : p1 - 1);
            int64_t lo = 0, hi = E.n1;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (E.offs[mid] <= p) lo = mid; else hi = mid;
            }
            s_r[threadIdx.x] = lo;
        }
        __syncthreads();
        const int64_t ia = s_r[0], ib = s_r[1];
        const bool cached = ib - ia < MJ_CACHE; // the usual case: one coalesced read of the slice's offsets, searches in shared memory
        if (cached) {
            for (int64_t k = threadIdx.x; k <= ib - ia; k += 256) {
                const unsigned long long o = E.offs[ia + k];
                s_off[k] = o <= (unsigned long long)p0 ? 0u : (uint32_t)min(o - (unsigned long long)p0, 0xffffffffull);
            }
            __syncthreads();
        }
        for (int64_t p = p0 + threadIdx.x; p < p1; p += 256) {
            int64_t lo = ia, hi = ib + 1;
            if (cached) {
                const uint32_t rel = (uint32_t)(p - p0);
                int l = 0, h = (int)(ib - ia) + 1; // s_off[0] = 0 <= rel
                while (h - l > 1) {
                    const int mid = (l + h) >> 1;
                    if (s_off[mid] <= rel) l = mid; else h = mid;
                }
                lo = ia + l;
            } else {
                while (hi - lo > 1) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (E.offs[mid] <= (unsigned long long)p) lo = mid; else hi = mid;
                }
            }
            const int64_t i = lo, j = p - (int64_t)E.offs[i];
            join_emit(E.C, p, E.rid1 ? (int64_t)E.rid1[i] : i, (int64_t)E.rid2[(int64_t)E.lb[i] + j]);
        }
        __syncthreads();
    }
}

// ---- hash join (order = 0) ----
constexpr int HJ_T = 256, HJ_I = 4, HJ_TILE = HJ_T * HJ_I;

// multiset build: every row gets its own entry (duplicates share a probe sequence)
// *dup is raised when two build rows carry the same 4-byte key (the later one meets the earlier one's entry on its way
// to a free slot): a table WITHOUT duplicates lets a probe stop at its first match instead of walking on to the next
// empty slot.  8-byte keys are stored after their entry is claimed, so they cannot be compared here: a second launch
// (hk_hj_dup8_kernel) looks every build row up in the finished table.
template <int KW>
__global__ void __launch_bounds__(256) hk_hj_build_kernel(const void *key_col, int64_t n, void *htab, unsigned long long hmask,
                                                          unsigned int *dup) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long *t = reinterpret_cast<unsigned long long *>(htab);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if constexpr (KW == 4) {
            const uint32_t key = reinterpret_cast<const uint32_t *>(key_col)[i];
            const unsigned long long packed = (unsigned long long)key | ((unsigned long long)(uint32_t)(i + 1) << 32);
            unsigned long long h = hk_hash_key<4>(key) & hmask;
            while (true) {
                const unsigned long long old = atomicCAS(t + h, 0ull, packed);
                if (old == 0ull) break;
                if ((uint32_t)old == key) *dup = 1u;
                h = (h + 1) & hmask;
            }
        } else {
            const unsigned long long key = reinterpret_cast<const unsigned long long *>(key_col)[i];
            unsigned long long h = hk_hash_key<8>(key) & hmask;
            while (atomicCAS(t + 2 * h + 1, 0ull, (unsigned long long)(uint32_t)(i + 1)) != 0ull) h = (h + 1) & hmask;
            t[2 * h] = key;
        }
    }
}

// 8-byte keys: a build row whose cluster holds another row with the same key raises *dup
__global__ void __launch_bounds__(256) hk_hj_dup8_kernel(const void *key_col, int64_t n, const void *htab, unsigned long long hmask,
                                                         unsigned int *dup) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const ulonglong2 *t = reinterpret_cast<const ulonglong2 *>(htab);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = reinterpret_cast<const unsigned long long *>(key_col)[i];
        unsigned long long h = hk_hash_key<8>(key) & hmask;
        while (true) {
            const ulonglong2 e = __ldg(t + h);
            if ((uint32_t)e.y == 0u) break;
            if (e.x == key && (uint32_t)e.y != (uint32_t)(i + 1)) {
                *dup = 1u;
                break;
            }
            h = (h + 1) & hmask;
        }
    }
}

// entry type of the multiset table and its first-probe load: the first probes of a thread's rows are all issued before
// any of them is examined (independent loads in flight), only collisions walk on
template <int KW> struct HjEntry;
template <> struct HjEntry<4> { using T = uint2; };
template <> struct HjEntry<8> { using T = ulonglong2; };

template <int KW>
__device__ __forceinline__ typename HjEntry<KW>::T hj_first(const void *htab, unsigned long long hmask, typename JRaw<KW>::T key,
                                                            unsigned long long *h) {
    *h = hk_hash_key<KW>(key) & hmask;
    return __ldg(reinterpret_cast<const typename HjEntry<KW>::T *>(htab) + *h);
}

// visits the build rows matching `key`, starting from the preloaded first entry; F(row) for each
template <int KW, typename F>
__device__ __forceinline__ void hj_for_matches(const void *htab, unsigned long long hmask, typename JRaw<KW>::T key,
                                               typename HjEntry<KW>::T e, unsigned long long h, F f, bool first_only = false) {
    const typename HjEntry<KW>::T *t = reinterpret_cast<const typename HjEntry<KW>::T *>(htab);
    while (true) {
        if constexpr (KW == 4) {
            if (e.y == 0u) return;
            if (e.x == key) {
                f((int64_t)e.y - 1);
                if (first_only) return;
            }
        } else {
            if ((uint32_t)e.y == 0u) return;
            if (e.x == key) {
                f((int64_t)(uint32_t)e.y - 1);
                if (first_only) return;
            }
        }
        h = (h + 1) & hmask;
        e = __ldg(t + h);
    }
}

struct HjParams {
    const void *k1;      // probe keys, table row order
    int64_t n1;
    const void *htab;
    unsigned long long hmask;
    unsigned long long *tile_base; // [num_tiles + 1] exclusive output offset of every probe tile
    uint64_t *state;
    unsigned long long *ticket;
    int64_t num_tiles;
    const unsigned int *dup; // 0: the build keys are unique
    JoinCols C;
};

template <int KW>
__global__ void __launch_bounds__(HJ_T) hk_hj_count_kernel(const __grid_constant__ HjParams P) {
    using KT = typename JRaw<KW>::T;
    __shared__ unsigned long long s_w[HJ_T / 32];
    __shared__ long long s_tile;
    const KT *k1 = reinterpret_cast<const KT *>(P.k1);
    const bool unique = *P.dup == 0u;
    while (true) {
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(P.ticket, 1ull);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.num_tiles) break;
        const int64_t i0 = tile * HJ_TILE;
        unsigned long long mine = 0;
        KT key[HJ_I];
        typename HjEntry<KW>::T e0[HJ_I];
        unsigned long long h0[HJ_I];
#pragma unroll
        for (int e = 0; e < HJ_I; e++) {
            const int64_t i = i0 + e * HJ_T + threadIdx.x;
            key[e] = i < P.n1 ? k1[i] : (KT)0;
        }
#pragma unroll
        for (int e = 0; e < HJ_I; e++) e0[e] = hj_first<KW>(P.htab, P.hmask, key[e], &h0[e]);
#pragma unroll
        for (int e = 0; e < HJ_I; e++) {
            const int64_t i = i0 + e * HJ_T + threadIdx.x;
            if (i < P.n1) hj_for_matches<KW>(P.htab, P.hmask, key[e], e0[e], h0[e], [&](int64_t) { mine++; }, unique);
        }
        unsigned long long total;
        mj_block_excl_scan(mine, s_w, &total);
        if (threadIdx.x < 32) {
            const unsigned long long ex = hk_lookback_u64(P.state, tile, total);
            if (threadIdx.x == 0) {
                P.tile_base[tile] = ex;
                if (tile == P.num_tiles - 1) P.tile_base[P.num_tiles] = ex + total;
            }
        }
        __syncthreads();
    }
}

// Unique build keys (the duplicate flag stayed down): every probe row has 0 or 1 match, so count, look-back and emit fit
// in ONE pass over the probe side — the table is walked once per row instead of twice, and the host allocates the result
// for n1 rows up front (its row count is set afterwards).  Rows are warp-striped like in hk_hj_expand_kernel, so the
// result keeps probe-row order and the stores coalesce.
template <int KW>
__global__ void __launch_bounds__(HJ_T) hk_hj_unique_kernel(const __grid_constant__ HjParams P) {
    using KT = typename JRaw<KW>::T;
    __shared__ unsigned long long s_w[HJ_T / 32];
    __shared__ long long s_tile;
    __shared__ unsigned long long s_base;
    const KT *k1 = reinterpret_cast<const KT *>(P.k1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    while (true) {
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(P.ticket, 1ull); // tiles in order: the look-back never waits on a tile that has not started
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.num_tiles) break;
        const int64_t i0 = tile * HJ_TILE + (int64_t)warp * (32 * HJ_I);
        KT key[HJ_I];
        typename HjEntry<KW>::T e0[HJ_I];
        unsigned long long h0[HJ_I];
        uint32_t m0[HJ_I], off[HJ_I], hit = 0, wsum = 0;
#pragma unroll
        for (int e = 0; e < HJ_I; e++) {
            const int64_t i = i0 + e * 32 + lane;
            key[e] = i < P.n1 ? k1[i] : (KT)0;
        }
#pragma unroll
        for (int e = 0; e < HJ_I; e++) e0[e] = hj_first<KW>(P.htab, P.hmask, key[e], &h0[e]);
#pragma unroll
        for (int e = 0; e < HJ_I; e++) {
            const int64_t i = i0 + e * 32 + lane;
            bool found = false;
            m0[e] = 0;
            if (i < P.n1)
                hj_for_matches<KW>(P.htab, P.hmask, key[e], e0[e], h0[e], [&](int64_t r2) {
                    m0[e] = (uint32_t)r2;
                    found = true;
                }, true);
            const uint32_t b = __ballot_sync(HK_FULL_MASK, found);
            if (found) hit |= 1u << e;
            off[e] = wsum + (uint32_t)__popc(b & lt_mask);
            wsum += (uint32_t)__popc(b);
        }
        if (lane == 0) s_w[warp] = wsum;
        __syncthreads();
        unsigned long long wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < HJ_T / 32; w++) {
            const unsigned long long x = s_w[w];
            if (w < warp) wbase += x;
            total += x;
        }
        if (threadIdx.x < 32) {
            const unsigned long long ex = hk_lookback_u64(P.state, tile, total);
            if (threadIdx.x == 0) {
                s_base = ex;
                if (tile == P.num_tiles - 1) P.tile_base[P.num_tiles] = ex + total;
            }
        }
        __syncthreads();
        wbase += s_base;
#pragma unroll
        for (int e = 0; e < HJ_I; e++)
            if (hit & (1u << e)) join_emit(P.C, (int64_t)(wbase + off[e]), i0 + e * 32 + lane, (int64_t)m0[e]);
    }
}

template <int KW>
__global__ void __launch_bounds__(HJ_T) hk_hj_expand_kernel(const __grid_constant__ HjParams P) {
    using KT = typename JRaw<KW>::T;
    __shared__ unsigned long long s_w[HJ_T / 32];
    const KT *k1 = reinterpret_cast<const KT *>(P.k1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool unique = *P.dup == 0u;
    for (int64_t tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        // a warp owns 128 consecutive probe rows, striped over its lanes (row = warp range + e * 32 + lane): loads and —
        // with one match per row — stores coalesce, and the result stays grouped by probe row in row order
        const int64_t i0 = tile * HJ_TILE + (int64_t)warp * (32 * HJ_I);
        KT key[HJ_I];
        typename HjEntry<KW>::T e0[HJ_I];
        unsigned long long h0[HJ_I];
        unsigned long long cnt[HJ_I], off[HJ_I], wsum = 0;
        uint32_t m0[HJ_I];
#pragma unroll
        for (int e = 0; e < HJ_I; e++) {
            const int64_t i = i0 + e * 32 + lane;
            key[e] = i < P.n1 ? k1[i] : (KT)0;
        }
#pragma unroll
        for (int e = 0; e < HJ_I; e++) e0[e] = hj_first<KW>(P.htab, P.hmask, key[e], &h0[e]);
#pragma unroll
        for (int e = 0; e < HJ_I; e++) {
            const int64_t i = i0 + e * 32 + lane;
            cnt[e] = 0;
            m0[e] = 0;
            if (i < P.n1)
                hj_for_matches<KW>(P.htab, P.hmask, key[e], e0[e], h0[e], [&](int64_t r2) {
                    if (cnt[e]++ == 0) m0[e] = (uint32_t)r2;
                }, unique);
            unsigned long long inc = cnt[e]; // inclusive warp scan in lane order
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(HK_FULL_MASK, inc, o);
                if (lane >= o) inc += t;
            }
            off[e] = wsum + inc - cnt[e];
            wsum += __shfl_sync(HK_FULL_MASK, inc, 31);
        }
        if (lane == 0) s_w[warp] = wsum;
        __syncthreads();
        unsigned long long wbase = P.tile_base[tile];
        for (int w = 0; w < warp; w++) wbase += s_w[w];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < HJ_I; e++) {
            const int64_t i = i0 + e * 32 + lane;
            if (i < P.n1 && cnt[e]) {
                unsigned long long run = wbase + off[e];
                if (cnt[e] == 1) { // the common case (a key of the build side occurs once): no second walk
                    join_emit(P.C, (int64_t)run, i, (int64_t)m0[e]);
                } else {
                    hj_for_matches<KW>(P.htab, P.hmask, key[e], e0[e], h0[e], [&](int64_t r2) {
                        join_emit(P.C, (int64_t)run, i, r2);
                        run++;
                    });
                }
            }
        }
    }
}

int fill_join_cols(hark_ctx *ctx, JoinCols &C, hark_table *t, const hark_table *db1, const hark_table *db2, const int32_t *cols1, int64_t l,
                   const int32_t *cols2, int64_t k) {
    C.l = (int)l;
    C.k = (int)k;
    for (int64_t j = 0; j < l; j++) {
        C.src1[j] = db1->cols[cols1[j]].ptr;
        C.w1[j] = hk_dtype_size(db1->cols[cols1[j]].dtype);
    }
    for (int64_t j = 0; j < k; j++) {
        C.src2[j] = db2->cols[cols2[j]].ptr;
        C.w2[j] = hk_dtype_size(db2->cols[cols2[j]].dtype);
    }
    for (int64_t j = 0; j < l + k; j++) C.dst[j] = t->cols[j].ptr;
    (void)ctx;
    return HARK_OK;
}

// (key, row id) of one side, sorted by key in `dtype` order, stable
// The left side with its projected columns CARRIED through the sort instead of a row id: the result of the ordered join
// is in key order, so a row id would cost one 64-byte DRAM access per projected value in the expand (ncu: 37 GB read for a
// 1.6 GB result at 1.34e8 rows); carried columns are read at the sorted position, coalesced.  carried[j] = sorted array
// of db->cols[cols[j]] (the key column maps to the sorted keys).
constexpr int MJ_CARRY_MAX = 3; // distinct non-key projected columns worth carrying (each adds its width to every pass)
int sort_side_carry(hark_ctx *ctx, Bufs &bufs, const hark_table *db, int32_t col, int32_t dtype, const int32_t *cols, int64_t l,
                    void **keys_out, std::vector<const void *> &carried) {
    const int64_t n = db->n;
    std::vector<hk_sort_array> arrays(1);
    arrays[0].in = db->cols[col].ptr;
    arrays[0].width = hk_dtype_size(dtype);
    std::vector<int> slot((size_t)db->cols.size(), -1);
    slot[(size_t)col] = 0;
    for (int64_t j = 0; j < l; j++) {
        const int c = cols[j];
        if (slot[(size_t)c] >= 0) continue;
        hk_sort_array a;
        a.in = db->cols[c].ptr;
        a.width = hk_dtype_size(db->cols[c].dtype);
        slot[(size_t)c] = (int)arrays.size();
        arrays.push_back(a);
    }
    std::vector<hk_sort_keyspec> keys{hk_sort_keyspec{0, dtype, 0}};
    HK_TRY(hk_radix_sort(ctx, n, keys, arrays, 0, nullptr, nullptr));
    for (auto &a : arrays) bufs.adopt(a.result);
    *keys_out = arrays[0].result;
    carried.assign((size_t)l, nullptr);
    for (int64_t j = 0; j < l; j++) carried[(size_t)j] = arrays[(size_t)slot[(size_t)cols[j]]].result;
    return HARK_OK;
}

int sort_side_typed(hark_ctx *ctx, Bufs &bufs, const hark_table *db, int32_t col, int32_t dtype, void **keys_out, void **rid_out) {
    const int64_t n = db->n;
    void *iota = nullptr;
    HK_TRY(bufs.alloc(&iota, sizeof(uint32_t) * (size_t)std::max<int64_t>(n, 1)));
    HK_TRY(hk_iota(ctx, iota, n, 4));
    std::vector<hk_sort_array> arrays(2);
    arrays[0].in = db->cols[col].ptr;
    arrays[0].width = hk_dtype_size(dtype);
    arrays[1].in = iota;
    arrays[1].width = 4;
    std::vector<hk_sort_keyspec> keys{hk_sort_keyspec{0, dtype, 0}};
    HK_TRY(hk_radix_sort(ctx, n, keys, arrays, 0, nullptr, nullptr));
    bufs.adopt(arrays[0].result);
    bufs.adopt(arrays[1].result);
    *keys_out = arrays[0].result;
    *rid_out = arrays[1].result;
    return HARK_OK;
}

} // namespace

int hk_join_ex(hark_ctx *ctx, hark_table **out, const hark_table *db1, const hark_table *db2, int32_t col1, int32_t col2,
               const int32_t *cols1, int64_t l, const int32_t *cols2, int64_t k, int32_t order, bool pinned_u32) {
    const int64_t n1 = db1->n, n2 = db2->n;
    const int64_t m1 = (int64_t)db1->cols.size(), m2 = (int64_t)db2->cols.size();
    HK_ARG(ctx, l <= JMAXC && k <= JMAXC, "join: at most 16 projected columns per side");
    // join.fut:55-56 slices db1[:,col1] / db2[:,col2] whenever the table has rows
    if (n1 > 0 || !pinned_u32) HK_ARG(ctx, col1 >= 0 && col1 < m1, "join: col1 out of bounds");
    if (n2 > 0 || !pinned_u32) HK_ARG(ctx, col2 >= 0 && col2 < m2, "join: col2 out of bounds");
    int32_t kdt = HARK_U32;
    if (pinned_u32) {
        for (auto *t : {db1, db2})
            for (auto &c : t->cols)
                HK_ARG(ctx, c.dtype == HARK_I32 || c.dtype == HARK_U32, "join: the reference entry takes u32 tables (use hark_entry_join_ex for typed tables)");
    } else {
        kdt = db1->cols[col1].dtype;
        HK_ARG(ctx, hk_dtype_int(kdt) && db2->cols[col2].dtype == kdt, "join_ex: both key columns must have the same integer dtype");
        // the projected columns are part of the call, whether or not any row comes out
        for (int64_t j = 0; j < l; j++) HK_ARG(ctx, cols1[j] >= 0 && cols1[j] < m1, "join: cols1 index out of bounds");
        for (int64_t j = 0; j < k; j++) HK_ARG(ctx, cols2[j] >= 0 && cols2[j] < m2, "join: cols2 index out of bounds");
    }
    const int kw = hk_dtype_size(kdt);
    ctx->entry_begin();
    std::vector<int32_t> odt((size_t)(l + k), HARK_U32);
    if (!pinned_u32) {
        for (int64_t j = 0; j < l; j++) odt[(size_t)j] = db1->cols[cols1[j]].dtype;
        for (int64_t j = 0; j < k; j++) odt[(size_t)(l + j)] = db2->cols[cols2[j]].dtype;
    }
    if (n1 == 0 || n2 == 0) {
        HK_TRY(hk_table_alloc(ctx, out, 0, 0, odt.data(), l + k));
        ctx->entry_end(0, n1 + n2, 0);
        return HARK_OK;
    }
    HK_ARG(ctx, n2 < 0xffffffffll, "join: more than 2^32-2 rows in db2 is not supported");
    Bufs bufs(ctx);
    auto check_proj = [&]() -> int { // join.fut:69-73 index rows of both tables with the projected columns (only when rows come out)
        for (int64_t j = 0; j < l; j++) HK_ARG(ctx, cols1[j] >= 0 && cols1[j] < m1, "join: cols1 index out of bounds");
        for (int64_t j = 0; j < k; j++) HK_ARG(ctx, cols2[j] >= 0 && cols2[j] < m2, "join: cols2 index out of bounds");
        return HARK_OK;
    };
    int64_t rowbytes = 0;
    hark_table *t = nullptr;
    int64_t P = 0;

    if (order != 0) {
        // ---- sort both sides, merge ----
        HK_ARG(ctx, n1 < 0xffffffffll, "join (ordered): more than 2^32-2 rows in db1 is not supported; use order = 0");
        void *k1 = nullptr, *r1 = nullptr, *k2 = nullptr, *r2 = nullptr;
        std::vector<const void *> carried1;
        bool carry = ctx->opt("join.carry", 1) != 0 && l > 0;
        {
            int distinct = 0;
            std::vector<char> seen((size_t)m1, 0);
            for (int64_t j = 0; j < l && carry; j++) {
                if (cols1[j] < 0 || cols1[j] >= m1) carry = false; // reported later, and only if rows come out (join.fut:69-73)
                else if (cols1[j] != col1 && !seen[(size_t)cols1[j]]) {
                    seen[(size_t)cols1[j]] = 1;
                    distinct++;
                }
            }
            if (distinct > MJ_CARRY_MAX) carry = false;
        }
        if (carry) HK_TRY(sort_side_carry(ctx, bufs, db1, col1, kdt, cols1, l, &k1, carried1));
        else HK_TRY(sort_side_typed(ctx, bufs, db1, col1, kdt, &k1, &r1));
        HK_TRY(sort_side_typed(ctx, bufs, db2, col2, kdt, &k2, &r2));
        ctx->counters["join.last_carry"] = carry ? 1 : 0;
        MjParams M;
        memset(&M, 0, sizeof M);
        M.k1 = k1;
        M.k2 = k2;
        M.n1 = n1;
        M.n2 = n2;
        M.dtype = kdt;
        M.num_tiles = (n1 + MJ_TILE - 1) / MJ_TILE;
        HK_TRY(bufs.alloc((void **)&M.lb, sizeof(uint32_t) * (size_t)n1));
        HK_TRY(bufs.alloc((void **)&M.offs, sizeof(unsigned long long) * (size_t)(n1 + 1)));
        HK_TRY(bufs.alloc((void **)&M.state, sizeof(uint64_t) * (size_t)(M.num_tiles + 1)));
        HK_CUDA(ctx, cudaMemsetAsync(M.state, 0, sizeof(uint64_t) * (size_t)(M.num_tiles + 1), ctx->stream));
        M.ticket = (unsigned long long *)(M.state + M.num_tiles);
        long long *ranges = nullptr;
        HK_TRY(bufs.alloc((void **)&ranges, sizeof(long long) * 2 * (size_t)M.num_tiles));
        M.ranges = ranges;
        ctx->kernel_begin();
        {
            const unsigned gr = (unsigned)((2 * M.num_tiles + 255) / 256);
            if (kw == 4) hk_mj_tile_ranges_kernel<4><<<gr, 256, 0, ctx->stream>>>(k1, k2, n1, n2, kdt, M.num_tiles, ranges);
            else hk_mj_tile_ranges_kernel<8><<<gr, 256, 0, ctx->stream>>>(k1, k2, n1, n2, kdt, M.num_tiles, ranges);
            HK_CHECK_LAUNCH(ctx);
        }
        const unsigned g = (unsigned)std::min<int64_t>(M.num_tiles, (int64_t)ctx->num_sms * 6);
        if (kw == 4) hk_mj_bounds_kernel<4><<<g, MJ_T, 0, ctx->stream>>>(M);
        else hk_mj_bounds_kernel<8><<<g, MJ_T, 0, ctx->stream>>>(M);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch(2);
        HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, M.offs + n1, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        P = (int64_t)ctx->h_scalars[0];
        if (P > 0) HK_TRY(check_proj());
        HK_TRY(hk_table_alloc(ctx, &t, P, P, odt.data(), l + k));
        if (P > 0) {
            MjExpandParams E;
            memset(&E, 0, sizeof E);
            E.P = P;
            E.n1 = n1;
            E.offs = M.offs;
            E.lb = M.lb;
            E.rid1 = (const uint32_t *)r1;
            E.rid2 = (const uint32_t *)r2;
            E.slice = 4096;
            fill_join_cols(ctx, E.C, t, db1, db2, cols1, l, cols2, k);
            if (carry)
                for (int64_t j = 0; j < l; j++) E.C.src1[j] = carried1[(size_t)j];
            const int64_t nslices = (P + E.slice - 1) / E.slice;
            hk_mj_expand_kernel<<<(unsigned)std::min<int64_t>(nslices, (int64_t)ctx->num_sms * 16), 256, 0, ctx->stream>>>(E);
            cudaError_t e = cudaGetLastError();
            ctx->count_launch();
            if (e != cudaSuccess) {
                hk_table_free(ctx, t);
                return ctx->fail(HARK_ERR_CUDA, std::string("join(expand): ") + cudaGetErrorString(e));
            }
        }
        ctx->kernel_end();
    } else {
        // ---- hash build on db2, probe with db1 ----
        uint64_t H = 1024;
        while (H < 2ull * (uint64_t)n2) H <<= 1;
        const size_t esz = kw == 4 ? 8 : 16;
        void *tab = nullptr;
        HK_TRY(bufs.alloc(&tab, (size_t)H * esz));
        HK_CUDA(ctx, cudaMemsetAsync(tab, 0, (size_t)H * esz, ctx->stream));
        unsigned int *dup = nullptr;
        HK_TRY(bufs.alloc((void **)&dup, sizeof(unsigned int)));
        HK_CUDA(ctx, cudaMemsetAsync(dup, 0, sizeof(unsigned int), ctx->stream));
        ctx->kernel_begin();
        if (kw == 4) hk_hj_build_kernel<4><<<grid_for(ctx, n2), 256, 0, ctx->stream>>>(db2->cols[col2].ptr, n2, tab, H - 1, dup);
        else hk_hj_build_kernel<8><<<grid_for(ctx, n2), 256, 0, ctx->stream>>>(db2->cols[col2].ptr, n2, tab, H - 1, dup);
        HK_CHECK_LAUNCH(ctx);
        if (kw == 8) {
            hk_hj_dup8_kernel<<<grid_for(ctx, n2), 256, 0, ctx->stream>>>(db2->cols[col2].ptr, n2, tab, H - 1, dup);
            HK_CHECK_LAUNCH(ctx);
            ctx->count_launch();
        }
        HjParams J;
        memset(&J, 0, sizeof J);
        J.dup = dup;
        J.k1 = db1->cols[col1].ptr;
        J.n1 = n1;
        J.htab = tab;
        J.hmask = H - 1;
        J.num_tiles = (n1 + HJ_TILE - 1) / HJ_TILE;
        HK_TRY(bufs.alloc((void **)&J.tile_base, sizeof(unsigned long long) * (size_t)(J.num_tiles + 1)));
        HK_TRY(bufs.alloc((void **)&J.state, sizeof(uint64_t) * (size_t)(J.num_tiles + 1)));
        HK_CUDA(ctx, cudaMemsetAsync(J.state, 0, sizeof(uint64_t) * (size_t)(J.num_tiles + 1), ctx->stream));
        J.ticket = (unsigned long long *)(J.state + J.num_tiles);
        const unsigned g = (unsigned)std::min<int64_t>(J.num_tiles, (int64_t)ctx->num_sms * 8);
        // the duplicate flag decides the plan: unique build keys -> one fused pass into a result allocated for n1 rows
        HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, dup, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const bool unique = *reinterpret_cast<const unsigned int *>(ctx->h_scalars) == 0u && ctx->opt("join.hash_one_pass", 1) != 0;
        ctx->counters["join.last_one_pass"] = unique ? 1 : 0;
        if (unique) {
            HK_TRY(check_proj());
            HK_TRY(hk_table_alloc(ctx, &t, n1, n1, odt.data(), l + k));
            fill_join_cols(ctx, J.C, t, db1, db2, cols1, l, cols2, k);
            if (kw == 4) hk_hj_unique_kernel<4><<<g, HJ_T, 0, ctx->stream>>>(J);
            else hk_hj_unique_kernel<8><<<g, HJ_T, 0, ctx->stream>>>(J);
            cudaError_t e = cudaGetLastError();
            ctx->count_launch(2);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_scalars, J.tile_base + J.num_tiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) {
                hk_table_free(ctx, t);
                return ctx->fail(HARK_ERR_CUDA, std::string("join(hash, one pass): ") + cudaGetErrorString(e));
            }
            P = (int64_t)ctx->h_scalars[0];
            t->n = P; // cap stays n1
            ctx->kernel_end();
            for (auto &c : t->cols) rowbytes += hk_dtype_size(c.dtype);
            ctx->entry_end((int64_t)kw * (n1 + n2) + P * rowbytes * 2, n1 + n2, P);
            *out = t;
            return HARK_OK;
        }
        if (kw == 4) hk_hj_count_kernel<4><<<g, HJ_T, 0, ctx->stream>>>(J);
        else hk_hj_count_kernel<8><<<g, HJ_T, 0, ctx->stream>>>(J);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch(2);
        HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, J.tile_base + J.num_tiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        P = (int64_t)ctx->h_scalars[0];
        HK_TRY(hk_table_alloc(ctx, &t, P, P, odt.data(), l + k));
        if (P > 0) {
            fill_join_cols(ctx, J.C, t, db1, db2, cols1, l, cols2, k);
            if (kw == 4) hk_hj_expand_kernel<4><<<g, HJ_T, 0, ctx->stream>>>(J);
            else hk_hj_expand_kernel<8><<<g, HJ_T, 0, ctx->stream>>>(J);
            cudaError_t e = cudaGetLastError();
            ctx->count_launch();
            if (e != cudaSuccess) {
                hk_table_free(ctx, t);
                return ctx->fail(HARK_ERR_CUDA, std::string("join(hash expand): ") + cudaGetErrorString(e));
            }
        }
        ctx->kernel_end();
    }
    for (auto &c : t->cols) rowbytes += hk_dtype_size(c.dtype);
    ctx->entry_end((int64_t)kw * (n1 + n2) + P * rowbytes * 2, n1 + n2, P);
    *out = t;
    return HARK_OK;
}

int hk_join(hark_ctx *ctx, hark_table **out, const hark_table *db1, const hark_table *db2, int32_t col1, int32_t col2,
            const int32_t *cols1, int64_t l, const int32_t *cols2, int64_t k) {
    return hk_join_ex(ctx, out, db1, db2, col1, col2, cols1, l, cols2, k, 1, true);
}

namespace {

// The open-addressing table of a dimension's key column (see hk_hash_build_kernel).  payload_row: entries carry the
// dimension row instead of the group slot.  Fails with HARK_ERR_ARG when the key is not unique.
int build_hash_table(hark_ctx *ctx, Bufs &bufs, const hark_table *dim, int32_t pk_col, int32_t g_col, unsigned long long g_lo,
                     int payload_row, void **htab_out, uint64_t *hmask_out) {
    const int64_t nd = dim->n;
    const int kw = hk_dtype_size(dim->cols[pk_col].dtype);
    uint64_t H = 1024;
    while (H < 2ull * (uint64_t)nd) H <<= 1;
    const size_t esz = kw == 4 ? 8 : 16;
    void *tab = nullptr;
    unsigned int *dup = nullptr;
    HK_TRY(bufs.alloc(&tab, (size_t)H * esz));
    HK_TRY(bufs.alloc((void **)&dup, sizeof(unsigned int)));
    HK_CUDA(ctx, cudaMemsetAsync(tab, 0, (size_t)H * esz, ctx->stream));
    HK_CUDA(ctx, cudaMemsetAsync(dup, 0, sizeof(unsigned int), ctx->stream));
    const void *pk = dim->cols[pk_col].ptr, *g = dim->cols[g_col].ptr;
    const int g_dtype = dim->cols[g_col].dtype;
    const int64_t slice_bytes = ctx->opt("join.lut_slice_bytes", 16ll << 20);
    if (kw == 4 && !payload_row && hk_dtype_size(g_dtype) == 4 && (int64_t)(H * esz) > slice_bytes * 3 / 2 && nd >= ctx->opt("join.build_partition_min_rows", 1 << 16) &&
        ctx->opt("join.build_partition", 1) != 0) {
        // a table much larger than L2: bring the dimension rows into table-slice order first (K8t), then build slice by slice
        hk_part_spec ps;
        ps.dtype = dim->cols[pk_col].dtype;
        ps.base = 0;
        ps.span = 0;
        ps.shift = 0;
        while (((uint64_t)slice_bytes / esz) >> (ps.shift + 1)) ps.shift++;
        while (((H - 1) >> ps.shift) + 1 > 256) ps.shift++;
        ps.nbins = (int)(((H - 1) >> ps.shift) + 1);
        hk_tpart tp;
        const void *gv[1] = {g};
        HK_TRY(hk_tile_partition(ctx, nd, pk, 4, ps, H - 1, 1, gv, &tp));
        bufs.adopt(tp.rows);
        bufs.adopt(tp.dir);
        unsigned long long *ticket = nullptr;
        HK_TRY(bufs.alloc((void **)&ticket, sizeof(unsigned long long)));
        HK_CUDA(ctx, cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), ctx->stream));
        hk_hash_build_tiles_kernel<<<(unsigned)ctx->num_sms * 8, 256, 0, ctx->stream>>>(tp.rows, tp.dir, tp.num_tiles, tp.nbins, g_dtype, g_lo,
                                                                                      tab, H - 1, dup, ticket);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
    } else if (kw == 4) {
        hk_hash_build_kernel<4><<<grid_for(ctx, nd), 256, 0, ctx->stream>>>(pk, g, g_dtype, nd, g_lo, payload_row, tab, H - 1, dup);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
    } else {
        hk_hash_build_kernel<8><<<grid_for(ctx, nd), 256, 0, ctx->stream>>>(pk, g, g_dtype, nd, g_lo, payload_row, tab, H - 1, dup);
        HK_CHECK_LAUNCH(ctx);
        hk_hash_verify_kernel<<<grid_for(ctx, nd), 256, 0, ctx->stream>>>(pk, nd, tab, H - 1, dup);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch(2);
    }
    HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, dup, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (((unsigned int *)ctx->h_scalars)[0] != 0) return ctx->fail(HARK_ERR_ARG, "join_groupby: dim.pk is not unique");
    *htab_out = tab;
    *hmask_out = H - 1;
    return HARK_OK;
}

} // namespace

// SELECT d.g, agg(f.s...) FROM fact f JOIN dim d ON f.fk = d.pk GROUP BY d.g, dim.pk unique.
// Build side: a direct-address lookup over [pk_min, pk_max] when the keys are dense (span <= 16 x rows), else an
// open-addressing hash table ("join.build": 0 auto, 1 lookup, 2 hash).  Probe side: fused into K2 (probe + aggregate in
// one pass over the fact columns, partitioned by slices of the build structure so the slice being probed is
// L2-resident) when the group domain fits one shared-memory table; otherwise probe -> filter -> GROUP BY.
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
    const int32_t fk_dtype = fact->cols[fk_col].dtype, pk_dtype = dim->cols[pk_col].dtype;

    long long *d_mm = nullptr;
    HK_TRY(bufs.alloc((void **)&d_mm, 2 * sizeof(long long)));
    long long pk_min = 0, pk_span = 0;
    uint32_t *lut = nullptr;
    void *htab = nullptr;
    uint64_t hmask = 0;
    bool use_hash = false;
    if (nd > 0) {
        // ---- key range of the build side ----
        ((long long *)ctx->h_scalars)[0] = LLONG_MAX;
        ((long long *)ctx->h_scalars)[1] = LLONG_MIN;
        HK_CUDA(ctx, cudaMemcpyAsync(d_mm, ctx->h_scalars, 2 * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        hk_minmax_int_kernel<<<grid_for(ctx, nd), 256, 0, ctx->stream>>>(dim->cols[pk_col].ptr, pk_dtype, nd, d_mm);
        HK_CHECK_LAUNCH(ctx);
        ctx->count_launch();
        HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, d_mm, 2 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        pk_min = ((long long *)ctx->h_scalars)[0];
        const long long pk_max = ((long long *)ctx->h_scalars)[1];
        const unsigned long long span = (unsigned long long)pk_max - (unsigned long long)pk_min + 1ull;
        const bool dense_ok = span != 0 && span <= (unsigned long long)nd * 16ull + (1ull << 22);
        const int64_t build = ctx->opt("join.build", 0);
        use_hash = build == 2 || (build != 1 && !dense_ok);
        if (!use_hash && !dense_ok)
            return ctx->fail(HARK_ERR_UNSUPPORTED, "join_groupby: join.build=1 (direct-address lookup) needs dense keys (span <= 16 x rows)");
        if (use_hash) {
            // the probe hashes the raw bits of fk: both key columns must have one representation
            const bool same = fk_dtype == pk_dtype || (hk_dtype_size(fk_dtype) == 4 && hk_dtype_size(pk_dtype) == 4 && pk_min >= 0 &&
                                                       pk_max <= 0x7fffffffll);
            if (!same)
                return ctx->fail(HARK_ERR_UNSUPPORTED, "join_groupby (hash build): fk and pk must have the same dtype");
        }
        pk_span = (long long)span;
    }
    ctx->counters["join.last_build"] = nd == 0 ? 0 : (use_hash ? 2 : 1);

    // ---- K2 in lookup / hash mode: probe and aggregate in one pass over the fact columns, no materialised join ----
    if (nd > 0 && ctx->opt("groupby.impl", 0) != 1 && nf > 0 && c <= HK_DENSE_MAX_AGGS) {
        hk_dense_req rq;
        rq.n = nf;
        rq.key = fact->cols[fk_col].ptr;
        rq.key_dtype = fk_dtype;
        rq.out_key_dtype = g_dtype;
        rq.c = (int)c;
        bool eligible = true;
        std::vector<int> val_of_col((size_t)mf, -1);
        for (int64_t j = 0; j < c && eligible; j++) {
            int code = ops[j];
            if (code < HARK_AGG_PROD || code > HARK_AGG_SUM64) code = HARK_AGG_MIN;
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
                rq.val_cols[rq.nvals] = &fact->cols[col];
                rq.val_dtypes[rq.nvals] = fact->cols[col].dtype;
                rq.nvals++;
            }
            rq.agg_val[j] = val_of_col[col];
        }
        if (eligible) {
            HK_TRY(hk_column_minmax(ctx, dim->cols[g_col], nd, g_dtype, &rq.g_lo, &rq.g_hi));
            if (rq.g_hi - rq.g_lo < (1ull << 20)) {
                if (use_hash) {
                    HK_TRY(build_hash_table(ctx, bufs, dim, pk_col, g_col, (unsigned long long)rq.g_lo, 0, &htab, &hmask));
                    rq.htab = htab;
                    rq.hmask = hmask;
                } else {
                    HK_TRY(bufs.alloc((void **)&lut, sizeof(uint32_t) * (size_t)pk_span));
                    HK_CUDA(ctx, cudaMemsetAsync(lut, 0, sizeof(uint32_t) * (size_t)pk_span, ctx->stream));
                    unsigned long long *nz = (unsigned long long *)d_mm;
                    HK_CUDA(ctx, cudaMemsetAsync(nz, 0, sizeof(unsigned long long), ctx->stream));
                    // a lookup much larger than L2: bring the dimension rows into lookup-slice order first, so the
                    // scattered stores below complete whole sectors while the slice is still L2-resident
                    const void *pk_src = dim->cols[pk_col].ptr, *g_src = dim->cols[g_col].ptr;
                    const int pkw = hk_dtype_size(pk_dtype);
                    if ((uint64_t)pk_span * 4ull > (96ull << 20) && gw == 4) {
                        hk_part_spec ps;
                        ps.dtype = pk_dtype;
                        ps.base = pkw == 4 ? (uint64_t)(ps.dtype == HARK_U32 ? (uint32_t)pk_min : ((uint32_t)(int32_t)pk_min ^ 0x80000000u))
                                           : ((uint64_t)pk_min ^ 0x8000000000000000ull);
                        ps.span = (uint64_t)pk_span;
                        ps.shift = 22; // 4 Mi entries = 16 MB of lookup per bin
                        while ((((uint64_t)pk_span - 1) >> ps.shift) + 1 > 256) ps.shift++;
                        ps.nbins = (int)((((uint64_t)pk_span - 1) >> ps.shift) + 1);
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
                        pk_src, pk_dtype, g_src, g_dtype, nd, pk_min, (unsigned long long)rq.g_lo, lut);
                    HK_CHECK_LAUNCH(ctx);
                    hk_count_nonzero_kernel<<<grid_for(ctx, (int64_t)pk_span / 4 + 1), 256, 0, ctx->stream>>>(lut, (int64_t)pk_span, nz);
                    HK_CHECK_LAUNCH(ctx);
                    ctx->count_launch(2);
                    HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, nz, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
                    HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                    if ((int64_t)ctx->h_scalars[0] != nd) return ctx->fail(HARK_ERR_ARG, "join_groupby: dim.pk is not unique");
                    rq.lut = lut;
                    rq.pk_min = pk_min;
                    rq.pk_span = pk_span;
                }
                bool handled = false;
                hark_table *t = nullptr;
                HK_TRY(hk_dense_groupby(ctx, &t, rq, &handled));
                if (handled) {
                    int64_t alg = nf * hk_dtype_size(rq.key_dtype) + nf * 4 * rq.nvals;
                    alg += nd * (hk_dtype_size(pk_dtype) + gw);
                    ctx->entry_end(alg, nf + nd, t->n);
                    *out = t;
                    return HARK_OK;
                }
                htab = nullptr; // built with group slots: the materialising path below needs dimension rows
            }
        }
    }

    // ---- materialising path: row lookup / hash table, probe -> (group key, hit) per fact row ----
    if (nd > 0) {
        if (use_hash) {
            HK_TRY(build_hash_table(ctx, bufs, dim, pk_col, g_col, 0ull, 1, &htab, &hmask));
        } else {
            if (!lut) HK_TRY(bufs.alloc((void **)&lut, sizeof(uint32_t) * (size_t)pk_span));
            HK_CUDA(ctx, cudaMemsetAsync(lut, 0, sizeof(uint32_t) * (size_t)pk_span, ctx->stream));
            unsigned int *dup = (unsigned int *)(d_mm);
            HK_CUDA(ctx, cudaMemsetAsync(dup, 0, sizeof(unsigned int), ctx->stream));
            hk_lut_build_kernel<<<grid_for(ctx, nd), 256, 0, ctx->stream>>>(dim->cols[pk_col].ptr, pk_dtype, nd, pk_min, lut, dup);
            HK_CHECK_LAUNCH(ctx);
            ctx->count_launch();
            HK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, dup, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
            HK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (((unsigned int *)ctx->h_scalars)[0] != 0) return ctx->fail(HARK_ERR_ARG, "join_groupby: dim.pk is not unique");
        }
    }
    void *gkey = nullptr;
    uint32_t *hit = nullptr;
    HK_TRY(bufs.alloc(&gkey, (size_t)std::max<int64_t>(nf, 1) * gw));
    HK_TRY(bufs.alloc((void **)&hit, sizeof(uint32_t) * (size_t)std::max<int64_t>(nf, 1)));
    ctx->kernel_begin();
    if (nf > 0) {
        if (nd == 0) {
            HK_CUDA(ctx, cudaMemsetAsync(hit, 0, sizeof(uint32_t) * (size_t)nf, ctx->stream));
            HK_CUDA(ctx, cudaMemsetAsync(gkey, 0, (size_t)nf * gw, ctx->stream));
        } else if (use_hash) {
            if (hk_dtype_size(fk_dtype) == 4)
                hk_hash_probe_kernel<4><<<grid_for(ctx, nf, 16), 256, 0, ctx->stream>>>(fact->cols[fk_col].ptr, nf, htab, hmask,
                                                                                      dim->cols[g_col].ptr, gw, gkey, hit);
            else
                hk_hash_probe_kernel<8><<<grid_for(ctx, nf, 16), 256, 0, ctx->stream>>>(fact->cols[fk_col].ptr, nf, htab, hmask,
                                                                                      dim->cols[g_col].ptr, gw, gkey, hit);
            HK_CHECK_LAUNCH(ctx);
            ctx->count_launch();
        } else {
            hk_probe_kernel<<<grid_for(ctx, nf, 16), 256, 0, ctx->stream>>>(fact->cols[fk_col].ptr, fk_dtype, nf, pk_min, pk_span, lut,
                                                                           dim->cols[g_col].ptr, gw, gkey, hit);
            HK_CHECK_LAUNCH(ctx);
            ctx->count_launch();
        }
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
    hk_table_free(ctx, matched);
    if (rc != HARK_OK) return rc;
    int64_t alg = 0;
    alg += nf * hk_dtype_size(fk_dtype);
    for (size_t j = 2; j < tmp.cols.size(); j++) alg += nf * hk_dtype_size(tmp.cols[j].dtype);
    alg += nd * (hk_dtype_size(pk_dtype) + gw);
    ctx->entry_end(alg, nf + nd, res->n);
    *out = res;
    return HARK_OK;
}
