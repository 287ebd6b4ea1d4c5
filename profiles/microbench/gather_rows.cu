// Microbenchmark 4: random reads of whole rows of R bytes (warp-coalesced, 16 B per lane), 7 rows in flight per
// warp pass — the access pattern of the wide-row COBS query.  Reports GB/s of row bytes.
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
// each warp handles "windows": 7 random rows; lanes cover the row in rounds of 512 B
__global__ void __launch_bounds__(256) k_rows(const uint8_t* __restrict__ tab, uint64_t n_rows, uint32_t R, uint64_t n_win, uint32_t* out) {
    uint32_t acc = 0;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w = warp; w < n_win; w += nwarp) {
        const uint8_t* a[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) a[j] = tab + (mix(w * 7 + j) % n_rows) * R;
        for (uint32_t off = lane * 16; off < R; off += 512) {
            uint4 m = make_uint4(~0u, ~0u, ~0u, ~0u);
#pragma unroll
            for (int j = 0; j < 7; ++j) { uint4 v = ld16(a[j] + off); m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w; }
            acc += m.x ^ m.y ^ m.z ^ m.w;
        }
    }
    if (acc == 0x12345678) out[0] = acc;
}
int main() {
    const uint64_t bytes = 5ULL << 30;
    uint8_t* tab; uint32_t* out;
    cudaMalloc(&tab, bytes); cudaMalloc(&out, 4); cudaMemset(tab, 0xA5, bytes);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (uint32_t R : {128u, 256u, 512u, 1280u, 2048u, 4096u, 16384u})
        for (int ctas : {2, 4, 8, 32}) {
            uint64_t n_rows = bytes / R, n_win = (4ULL << 30) / (7ULL * R);
            k_rows<<<sm * ctas, 256>>>(tab, n_rows, R, n_win / 8, out);
            cudaEventRecord(a);
            k_rows<<<sm * ctas, 256>>>(tab, n_rows, R, n_win, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("R=%5u B grid=%2d ctas/sm: %8.2f ms  %7.1f GB/s\n", R, ctas, ms, n_win * 7.0 * R / ms / 1e6);
        }
    printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
