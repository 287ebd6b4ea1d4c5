// Microbenchmark: random 16-byte row gathers confined to a window of the table that moves slowly (the access pattern
// of k_bucket_fetch: all warps sweep one row range at a time), for window sizes from 1 MB to the whole table, alone
// and next to a streaming 16-byte write + 4-byte read per gather (plain and evict-first).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_window gather_window.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
__device__ __forceinline__ uint4 ld16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// STREAM: 0 none, 1 plain st/ld, 2 .cs st/ld
template <int STREAM>
__global__ void __launch_bounds__(256) k_window(const uint8_t* __restrict__ tab, uint64_t n_rows, uint64_t win_rows, uint64_t items_per_win,
                                                uint64_t n_items, const uint32_t* __restrict__ in, uint4* __restrict__ out, uint32_t* sink) {
    uint32_t acc = 0;
    // win_rows and items_per_win are powers of two (shifts and masks only: the loop must not be ALU-bound)
    const uint32_t wshift = 63 - __clzll(items_per_win), rshift = 63 - __clzll(win_rows);
    const uint64_t n_win = n_rows >> rshift;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; i0 < n_items; i0 += (uint64_t)gridDim.x * blockDim.x * 4) {
        uint4 v[4]; uint32_t r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint64_t i = i0 + q * 256;
            if (STREAM == 1) r[q] = in[i]; else if (STREAM == 2) asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(r[q]) : "l"(in + i)); else r[q] = 0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint64_t i = i0 + q * 256;
            uint64_t w = i >> wshift; w = w >= n_win ? w - n_win * (w / n_win) : w;
            uint64_t row = (w << rshift) + ((mix(i) + r[q]) & (win_rows - 1));
            v[q] = ld16(tab + row * 16);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint64_t i = i0 + q * 256;
            if (STREAM == 1) out[i] = v[q];
            else if (STREAM == 2) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out + i), "r"(v[q].x), "r"(v[q].y), "r"(v[q].z), "r"(v[q].w) : "memory");
            else acc += v[q].x ^ v[q].w;
        }
    }
    if (acc == 0x12345678) sink[0] = acc;
}

int main() {
    const uint64_t n_rows = 150000000ULL, n_items = 1ULL << 29;   // 2.4 GB table, 537 M gathers (8.6 GB of stream writes)
    uint8_t* tab; uint32_t* in; uint4* out; uint32_t* sink;
    cudaMalloc(&tab, n_rows * 16); cudaMalloc(&in, n_items * 4); cudaMalloc(&out, n_items * 16); cudaMalloc(&sink, 4);
    cudaMemset(tab, 0xA5, n_rows * 16); cudaMemset(in, 0, n_items * 4);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const uint64_t wins_mb[] = {1, 4, 16, 32, 64, 2048};
    for (int stream = 0; stream < 3; ++stream)
        for (uint64_t mb : wins_mb) {
            uint64_t win_rows = (mb << 20) / 16;
            // every row of a window is touched ~8 times per visit (the bench: 910 M probes per pass over 150 M rows = 6)
            uint64_t items_per_win = win_rows * 8;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (stream == 0) k_window<0><<<sm * 8, 256>>>(tab, n_rows, win_rows, items_per_win, n_items, in, out, sink);
                if (stream == 1) k_window<1><<<sm * 8, 256>>>(tab, n_rows, win_rows, items_per_win, n_items, in, out, sink);
                if (stream == 2) k_window<2><<<sm * 8, 256>>>(tab, n_rows, win_rows, items_per_win, n_items, in, out, sink);
                cudaEventRecord(b); cudaEventSynchronize(b);
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("stream=%d window=%5llu MB  %8.2f ms  %7.2f Ggather/s\n", stream, (unsigned long long)mb, ms, n_items / ms / 1e6);
        }
    printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
