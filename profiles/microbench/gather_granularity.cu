// Microbenchmark: random 16-byte row gathers from a table much larger than L2 (the access pattern of the
// narrow-row COBS query), under different load flavours and L2 fetch-granularity limits.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_granularity gather_granularity.cu
// Run:   ./gather_granularity [limit_bytes]   (limit applied with cudaDeviceSetLimit before any allocation)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
template <int MODE>
__device__ __forceinline__ uint4 ld16(const uint8_t* p) {
    uint4 v;
    if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 1) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 2) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 3) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 4) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 6) asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 7) asm volatile("ld.global.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 8) asm volatile("ld.global.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 9) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 5) asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int MODE, int PER>
__global__ void __launch_bounds__(256) k_gather(const uint8_t* __restrict__ tab, uint64_t n_rows, uint64_t n_items, uint32_t* out) {
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 m = make_uint4(~0u, ~0u, ~0u, ~0u);
        const uint8_t* a[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) a[j] = tab + (mix(i * PER + j) % n_rows) * 16;
#pragma unroll
        for (int j = 0; j < PER; ++j) { uint4 v = ld16<MODE>(a[j]); m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w; }
        acc += m.x ^ m.y ^ m.z ^ m.w;
    }
    if (acc == 0x12345678) out[0] = acc;
}
template <int MODE>
float run(const uint8_t* tab, uint64_t n_rows, uint64_t n_items, uint32_t* out, int grid) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_gather<MODE, 7><<<grid, 256>>>(tab, n_rows, n_items / 8, out);
    cudaEventRecord(a);
    k_gather<MODE, 7><<<grid, 256>>>(tab, n_rows, n_items, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}
int main(int argc, char** argv) {
    size_t lim = 0;
    if (argc > 1) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1]));
        printf("set limit %s -> %s\n", argv[1], cudaGetErrorString(e));
    }
    cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity = %zu\n", lim);
    const uint64_t n_rows = 150000001ULL, n_items = 1ULL << 27;   // 2.4 GB table, 134M items x 7 gathers
    uint8_t* tab; uint32_t* out;
    cudaMalloc(&tab, n_rows * 16); cudaMalloc(&out, 4);
    cudaMemset(tab, 0xA5, n_rows * 16);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    for (int occ : {4, 8}) {
        int grid = sm * occ;
        float ms[10] = {run<0>(tab, n_rows, n_items, out, grid), run<1>(tab, n_rows, n_items, out, grid), run<2>(tab, n_rows, n_items, out, grid),
                       run<3>(tab, n_rows, n_items, out, grid), run<4>(tab, n_rows, n_items, out, grid), run<5>(tab, n_rows, n_items, out, grid),
                       run<6>(tab, n_rows, n_items, out, grid), run<7>(tab, n_rows, n_items, out, grid), run<8>(tab, n_rows, n_items, out, grid),
                       run<9>(tab, n_rows, n_items, out, grid)};
        const char* names[10] = {"nc.L1::no_allocate", "plain", "cg", "cv", "L1::no_allocate (generic)", "lu", "L2::64B", "L2::128B", "L2::256B", "nc.L2::64B"};
        for (int m = 0; m < 10; ++m)
            printf("ctas/sm=%d %-34s %8.2f ms  %7.2f Ggather/s  %7.1f GB/s @32B-sector  %7.1f GB/s @128B-line\n", occ, names[m], ms[m],
                   n_items * 7 / ms[m] / 1e6, n_items * 7 * 32 / ms[m] / 1e6, n_items * 7 * 128 / ms[m] / 1e6);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    return 0;
}
