// Microbenchmark 3: the same 7-gather kernel issued as one launch vs split over 2-4 concurrent streams.
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
// REGS pads register use so that occupancy matches the real kernel (64 regs -> 4 CTAs/SM of 256)
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_g(const uint8_t* __restrict__ tab, uint64_t n_rows, uint64_t i_begin, uint64_t i_end, uint32_t* out) {
    uint32_t acc = 0;
    for (uint64_t i = i_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i_end; i += (uint64_t)gridDim.x * blockDim.x) {
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
int main() {
    const uint64_t n_rows = 150000001ULL, n_items = 1ULL << 28;
    uint8_t* tab; uint32_t* out;
    cudaMalloc(&tab, n_rows * 16); cudaMalloc(&out, 4); cudaMemset(tab, 0xA5, n_rows * 16);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    cudaStream_t st[8]; for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int ctas : {2, 4, 8, 16})
        for (int ns : {1, 2, 4}) {
            for (int chunks : {ns, 16}) {
                float best = 1e9;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaDeviceSynchronize();
                    cudaEventRecord(a, st[0]);
                    for (int s = 1; s < ns; ++s) cudaStreamWaitEvent(st[s], a);
                    for (int c = 0; c < chunks; ++c) {
                        uint64_t b0 = n_items * c / chunks, b1 = n_items * (c + 1) / chunks;
                        k_g<1><<<sm * ctas, 256, 0, st[c % ns]>>>(tab, n_rows, b0, b1, out);
                    }
                    cudaEvent_t e[8];
                    for (int s = 1; s < ns; ++s) { cudaEventCreate(&e[s]); cudaEventRecord(e[s], st[s]); cudaStreamWaitEvent(st[0], e[s]); }
                    cudaEventRecord(b, st[0]); cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
                }
                printf("grid=%2d ctas/sm streams=%d chunks=%2d: %8.2f ms %7.2f Ggather/s\n", ctas, ns, chunks, best, n_items * 7 / best / 1e6);
            }
        }
    printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
