// Microbenchmark 2: does deepening the DRAM queue without registers help random 16-byte gathers?
//   A  plain: 7 gathers per item into registers
//   B  prefetch.global.L2 of the next LAG items' rows, then the demand loads hit L2
//   C  cp.async (LDGSTS) 16 B into shared memory, NW items (7 rows each) in flight per thread
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_prefetch gather_prefetch.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
__device__ __forceinline__ uint4 ld16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void pf(const uint8_t* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

template <int LAG>
__global__ void __launch_bounds__(256) k_pf(const uint8_t* __restrict__ tab, uint64_t n_rows, uint64_t n_items, uint32_t* out) {
    uint32_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (LAG > 0) {
        for (int l = 0; l < LAG; ++l)
#pragma unroll
            for (int j = 0; j < 7; ++j) pf(tab + (mix((i0 + l * stride) * 7 + j) % n_rows) * 16);
    }
    for (uint64_t i = i0; i < n_items; i += stride) {
        if (LAG > 0) {
            uint64_t ip = i + LAG * stride;
#pragma unroll
            for (int j = 0; j < 7; ++j) pf(tab + (mix(ip * 7 + j) % n_rows) * 16);
        }
        uint4 m = make_uint4(~0u, ~0u, ~0u, ~0u);
        const uint8_t* a[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) a[j] = tab + (mix(i * 7 + j) % n_rows) * 16;
#pragma unroll
        for (int j = 0; j < 7; ++j) { uint4 v = ld16(a[j]); m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w; }
        acc += m.x ^ m.y ^ m.z ^ m.w;
    }
    if (acc == 0x12345678) out[0] = acc;
}

template <int NW, int NT>
__global__ void __launch_bounds__(NT) k_cpasync(const uint8_t* __restrict__ tab, uint64_t n_rows, uint64_t n_items, uint32_t* out) {
    extern __shared__ __align__(16) uint4 sm[];   // [NW][7][NT]
    uint32_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride * NW) {
#pragma unroll
        for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const uint8_t* a = tab + (mix((i + w * stride) * 7 + j) % n_rows) * 16;
                uint32_t s = (uint32_t)__cvta_generic_to_shared(&sm[(w * 7 + j) * NT + threadIdx.x]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(a));
            }
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint4 m = make_uint4(~0u, ~0u, ~0u, ~0u);
#pragma unroll
            for (int j = 0; j < 7; ++j) { uint4 v = sm[(w * 7 + j) * NT + threadIdx.x]; m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w; }
            acc += m.x ^ m.y ^ m.z ^ m.w;
        }
    }
    if (acc == 0x12345678) out[0] = acc;
}

template <class F>
float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(8); cudaEventRecord(a); f(1); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    const uint64_t n_rows = 150000001ULL, n_items = 1ULL << 27;
    uint8_t* tab; uint32_t* out;
    cudaMalloc(&tab, n_rows * 16 + 4096); cudaMalloc(&out, 4); cudaMemset(tab, 0xA5, n_rows * 16);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    auto rep = [&](const char* name, float ms) { printf("%-44s %8.2f ms %7.2f Ggather/s\n", name, ms, n_items * 7 / ms / 1e6); };
    rep("plain 4 ctas/sm", timeit([&](int d) { k_pf<0><<<sm * 4, 256>>>(tab, n_rows, n_items / d, out); }));
    rep("plain 8 ctas/sm", timeit([&](int d) { k_pf<0><<<sm * 8, 256>>>(tab, n_rows, n_items / d, out); }));
    rep("prefetch lag1 4 ctas/sm", timeit([&](int d) { k_pf<1><<<sm * 4, 256>>>(tab, n_rows, n_items / d, out); }));
    rep("prefetch lag2 4 ctas/sm", timeit([&](int d) { k_pf<2><<<sm * 4, 256>>>(tab, n_rows, n_items / d, out); }));
    rep("prefetch lag4 4 ctas/sm", timeit([&](int d) { k_pf<4><<<sm * 4, 256>>>(tab, n_rows, n_items / d, out); }));
    rep("prefetch lag8 4 ctas/sm", timeit([&](int d) { k_pf<8><<<sm * 4, 256>>>(tab, n_rows, n_items / d, out); }));
    rep("prefetch lag2 8 ctas/sm", timeit([&](int d) { k_pf<2><<<sm * 8, 256>>>(tab, n_rows, n_items / d, out); }));
    {
        auto k = k_cpasync<2, 128>; size_t s = 2 * 7 * 128 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s);
        int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 128, s);
        char nm[64]; snprintf(nm, 64, "cp.async NW=2 NT=128 occ=%d", occ);
        rep(nm, timeit([&](int d) { k<<<sm * occ, 128, s>>>(tab, n_rows, n_items / d, out); }));
    }
    {
        auto k = k_cpasync<2, 256>; size_t s = 2 * 7 * 256 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s);
        int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, s);
        char nm[64]; snprintf(nm, 64, "cp.async NW=2 NT=256 occ=%d", occ);
        rep(nm, timeit([&](int d) { k<<<sm * occ, 256, s>>>(tab, n_rows, n_items / d, out); }));
    }
    {
        auto k = k_cpasync<4, 128>; size_t s = 4 * 7 * 128 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s);
        int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 128, s);
        char nm[64]; snprintf(nm, 64, "cp.async NW=4 NT=128 occ=%d", occ);
        rep(nm, timeit([&](int d) { k<<<sm * occ, 128, s>>>(tab, n_rows, n_items / d, out); }));
    }
    {
        auto k = k_cpasync<1, 256>; size_t s = 1 * 7 * 256 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s);
        int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, s);
        char nm[64]; snprintf(nm, 64, "cp.async NW=1 NT=256 occ=%d", occ);
        rep(nm, timeit([&](int d) { k<<<sm * occ, 256, s>>>(tab, n_rows, n_items / d, out); }));
    }
    printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
