// sort.cuh — internal interface of the radix sort / partition / gather building blocks (sort.cu).
#pragma once
#include <vector>

#include "hark_internal.cuh"

#define HK_SORT_MAX_ARRAYS 20
#define HK_PEER_MAX 16 // GPUs a peer scatter can address (one NVSwitch domain)

struct hk_sort_keyspec {
    int array;     // index into the carried arrays
    int32_t dtype; // hark_dtype that defines the order (U32 for the reference-pinned paths)
    int32_t desc;
    // optional: min / max order key of the column when the caller already knows them (column statistics)
    bool have_range = false;
    uint64_t lo = 0, hi = 0;
};

struct hk_sort_array {
    const void *in = nullptr; // input (never written)
    int width = 4;            // 4 or 8 bytes
    void *buf[2] = {nullptr, nullptr};
    void *result = nullptr;   // out: fresh buffer from the context pool, owned by the caller
};

struct hk_sort_info {
    int passes = 0;
    int64_t bytes_moved = 0;
};

// keys: most significant first.  hash_nparts > 0: one partition pass by mix(keys[0]) % nparts instead of a sort;
// hash_counts[nparts] then receives the bucket sizes.
int hk_radix_sort(hark_ctx *ctx, int64_t n, const std::vector<hk_sort_keyspec> &keys, std::vector<hk_sort_array> &arrays,
                  int hash_nparts, int64_t *hash_counts, hk_sort_info *info);
int hk_iota(hark_ctx *ctx, void *out, int64_t n, int width);
int hk_gather(hark_ctx *ctx, void *dst, const void *src, int vwidth, const void *perm, int pwidth, int64_t n);
int hk_copy_bytes(hark_ctx *ctx, void *dst, const void *src, int64_t bytes);
int hk_peer_scatter_pass(hark_ctx *ctx, int64_t n, const void *digit, const void *const *cols, const int *widths, int ncols,
                         const unsigned long long *d_peer_out);
