// Microbenchmark: random 16-byte row gathers from an L2-resident 16 MB window through the TMA unit
// (cp.async.bulk global -> shared, one 16-byte copy per row, completion on an mbarrier) against the same gathers
// through the LSU (ld.global.nc.v4).  Question: does the bulk-copy path get around the L1TEX tag-stage limit
// (one lookup per gathered row, ~170 G rows/s chip-wide) that bounds k_bucket_fetch?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_tma gather_tma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int NT = 256, PER = 4;

__global__ void __launch_bounds__(NT) k_lsu(const uint8_t* __restrict__ tab, uint64_t win_mask, uint64_t n_items, uint32_t* sink) {
    uint32_t acc = 0;
    for (uint64_t i0 = (uint64_t)blockIdx.x * NT * PER + threadIdx.x; i0 < n_items; i0 += (uint64_t)gridDim.x * NT * PER) {
        uint4 v[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const uint8_t* p = tab + (mix(i0 + q * NT) & win_mask) * 16;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[q].x), "=r"(v[q].y), "=r"(v[q].z), "=r"(v[q].w) : "l"(p));
        }
#pragma unroll
        for (int q = 0; q < PER; ++q) acc += v[q].x ^ v[q].w;
    }
    if (acc == 0x12345678) sink[0] = acc;
}

__global__ void __launch_bounds__(NT) k_tma(const uint8_t* __restrict__ tab, uint64_t win_mask, uint64_t n_items, uint32_t* sink) {
    __shared__ __align__(16) uint4 buf[NT * PER];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_a), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    uint32_t acc = 0, phase = 0;
    for (uint64_t i0 = (uint64_t)blockIdx.x * NT * PER + threadIdx.x; i0 < n_items; i0 += (uint64_t)gridDim.x * NT * PER) {
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(NT * PER * 16) : "memory");
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const uint8_t* p = tab + (mix(i0 + q * NT) & win_mask) * 16;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                         :: "r"(smem_u32(&buf[q * NT + threadIdx.x])), "l"(p), "r"(bar_a) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
        phase ^= 1;
#pragma unroll
        for (int q = 0; q < PER; ++q) { uint4 v = buf[q * NT + threadIdx.x]; acc += v.x ^ v.w; }
        __syncthreads();   // buffer reuse
    }
    if (acc == 0x12345678) sink[0] = acc;
}

int main() {
    const uint64_t n_rows = 1ULL << 27, n_items = 1ULL << 29;   // 2 GB table
    uint8_t* tab; uint32_t* sink;
    cudaMalloc(&tab, n_rows * 16); cudaMalloc(&sink, 4);
    cudaMemset(tab, 0xA5, n_rows * 16);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (uint64_t mb : {1ULL, 16ULL, 2048ULL}) {
        const uint64_t win_mask = (mb << 20) / 16 - 1;
        for (int mode = 0; mode < 2; ++mode)
            for (int occ : {4, 8}) {
                float ms = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(a);
                    if (mode == 0) k_lsu<<<sm * occ, NT>>>(tab, win_mask, n_items, sink);
                    else k_tma<<<sm * occ, NT>>>(tab, win_mask, n_items, sink);
                    cudaEventRecord(b); cudaEventSynchronize(b);
                    cudaEventElapsedTime(&ms, a, b);
                }
                printf("window=%5llu MB  %s ctas/sm=%d  %8.2f ms  %7.2f Ggather/s\n", (unsigned long long)mb, mode ? "cp.async.bulk 16B" : "ld.global.nc.v4  ",
                       occ, ms, n_items / ms / 1e6);
            }
    }
    printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
