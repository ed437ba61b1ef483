// Microbenchmark: throughput of many SMALL cp.async.bulk shared->global copies (one per digit run of a radix-sort tile).
// Each CTA (512 threads) owns a 64 KB stage; per "tile" the first NRUN threads issue one bulk copy of RUNB bytes each to
// scattered, 16-byte aligned destinations; then commit + wait.  Prints GB/s per (RUNB, NRUN).  nvcc -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_store(void *g, const void *s, uint32_t bytes) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(g)), "r"(sa), "r"(bytes) : "memory");
}

template <int RUNB>
__global__ void __launch_bounds__(512, 2) k(unsigned char *out, size_t out_bytes, int tiles, int nrun) {
    extern __shared__ __align__(128) unsigned char st[];
    for (int i = threadIdx.x; i < 65536 / 4; i += 512) reinterpret_cast<uint32_t *>(st)[i] = i * 2654435761u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const size_t region = out_bytes / nrun;            // each run index writes into its own region (like a digit bin)
    for (int t = 0; t < tiles; t++) {
        const size_t tile_id = (size_t)blockIdx.x * tiles + t;
        if (threadIdx.x < nrun) {
            unsigned char *dst = out + (size_t)threadIdx.x * region + (tile_id * RUNB) % (region - RUNB);
            dst = (unsigned char *)((uintptr_t)dst & ~(uintptr_t)15);
            bulk_store(dst, st + threadIdx.x * RUNB, RUNB);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
    if (threadIdx.x < nrun) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int RUNB>
void run(unsigned char *d, size_t bytes, int nrun) {
    const int tiles = 2000, grid = 296;
    cudaFuncSetAttribute(k<RUNB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<RUNB><<<grid, 512, 65536>>>(d, bytes, 10, nrun);
    cudaEventRecord(a);
    k<RUNB><<<grid, 512, 65536>>>(d, bytes, tiles, nrun);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double gb = (double)grid * tiles * nrun * RUNB / 1e9;
    printf("run %4d B x %3d runs/tile: %8.3f ms  %8.1f GB/s  %8.1f M copies/s  (%s)\n", RUNB, nrun, ms, gb / (ms * 1e-3), (double)grid * tiles * nrun / ms / 1e3,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    size_t bytes = 8ull << 30;
    unsigned char *d;
    cudaMalloc(&d, bytes);
    run<64>(d, bytes, 256);
    run<128>(d, bytes, 256);
    run<256>(d, bytes, 256);
    run<512>(d, bytes, 128);
    run<1024>(d, bytes, 64);
    run<4096>(d, bytes, 16);
    run<16384>(d, bytes, 4);
    return 0;
}
