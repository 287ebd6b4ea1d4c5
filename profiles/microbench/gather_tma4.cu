// Microbenchmark (round 2): can the TMA unit serve the random 16-byte row gathers of k_bucket_fetch faster than the
// LSU?  Round 1 measured synchronous 16-byte cp.async.bulk copies at 72-81 G rows/s against 145-168 G/s through
// ld.global.nc.v4 (L1TEX: ~2 cycles per gathered line).  This one keeps a ring of mbarrier stages in flight per warp
// and adds Blackwell's tile::gather4 tensor copy (one instruction fetches 4 rows of a 2-D tensor by row index):
//
//   mode 0  lsu      ld.global.nc.v4, 4 gathers per lane in flight (the k_bucket_fetch loop)
//   mode 1  bulk16   cp.async.bulk 16 B per row, STAGES x 4 copies per lane in flight
//   mode 2  gather4  cp.async.bulk.tensor.2d.tile::gather4, one instruction per 4 rows, STAGES in flight per lane
//   mode 3  hybrid   per round: 4 rows per lane through the LSU + 4 rows per lane through gather4
//
// Rows come from a window of the table (16 MB = L2 resident like a bucket, or the whole 2 GB table = DRAM).
// Each mode runs in its own process (argv[1]) so that a faulting variant cannot poison the others.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_tma4 gather_tma4.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ uint4 ld16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

constexpr int NT = 128, NW = NT / 32, STAGES = 4;
constexpr int SLOT = 128;                         // bytes of shared memory per lane and stage (tensor copies want 128-byte alignment)

// rounds of 128 rows per warp: lane l owns rows 4l .. 4l+3 of the round
template <int MODE>
__global__ void __launch_bounds__(NT) k_gather(const uint8_t* __restrict__ tab, const __grid_constant__ CUtensorMap tmap,
                                               uint64_t win_mask, uint64_t n_rounds, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t s_buf[];          // [NW][STAGES][32][SLOT]
    __shared__ __align__(8) uint64_t s_bar[NW][STAGES];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&s_bar[warp][s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint64_t gw = (uint64_t)blockIdx.x * NW + warp, n_gw = (uint64_t)gridDim.x * NW;
    uint32_t acc = 0;

    auto row_of = [&](uint64_t round, int q) { return (uint32_t)(mix(round * 128 + lane * 4 + q) & win_mask); };
    auto issue = [&](uint64_t round, int s) {
        const uint32_t bar = smem_u32(&s_bar[warp][s]);
        const uint32_t dst = smem_u32(s_buf + ((size_t)(warp * STAGES + s) * 32 + lane) * SLOT);
        if (lane == 0) mbar_expect(bar, 32 * 64);
        __syncwarp();
        if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                             :: "r"(dst + q * 16), "l"(tab + (uint64_t)row_of(round, q) * 16), "r"(bar) : "memory");
        } else {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                         " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                         :: "r"(dst), "l"(&tmap), "r"(bar), "r"(0), "r"(row_of(round, 0)), "r"(row_of(round, 1)),
                            "r"(row_of(round, 2)), "r"(row_of(round, 3)) : "memory");
        }
    };

    if (MODE == 0) {
        for (uint64_t r = gw; r < n_rounds; r += n_gw) {
            uint4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = ld16(tab + (uint64_t)row_of(r, q) * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += v[q].x ^ v[q].w;
        }
    } else {
        // MODE 3: every other round goes through the LSU while the ring is in flight
        const uint64_t stride = (MODE == 3) ? 2 * n_gw : n_gw;
        uint64_t next = gw;
        uint32_t phase = 0;
        int n_inflight = 0;
        for (int s = 0; s < STAGES && next < n_rounds; ++s, next += stride, ++n_inflight) issue(next, s);
        int s = 0;
        uint64_t lsu_round = gw + n_gw;
        while (n_inflight > 0) {
            if (MODE == 3 && lsu_round < n_rounds) {
                uint4 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = ld16(tab + (uint64_t)row_of(lsu_round, q) * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc += v[q].x ^ v[q].w;
                lsu_round += stride;
            }
            mbar_wait(smem_u32(&s_bar[warp][s]), phase);
            const uint4* src = reinterpret_cast<const uint4*>(s_buf + ((size_t)(warp * STAGES + s) * 32 + lane) * SLOT);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const uint4 v = src[q]; acc += v.x ^ v.w; }
            __syncwarp();
            --n_inflight;
            if (next < n_rounds) { issue(next, s); next += stride; ++n_inflight; }
            if (++s == STAGES) { s = 0; phase ^= 1; }
        }
        if (MODE == 3)
            for (; lsu_round < n_rounds; lsu_round += stride) {
                uint4 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = ld16(tab + (uint64_t)row_of(lsu_round, q) * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc += v[q].x ^ v[q].w;
            }
    }
    if (acc == 0x12345678) sink[0] = acc;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int box_rows = argc > 2 ? atoi(argv[2]) : 1;       // tensor-map box rows for gather4: 1 or 4 (both tried)
    const uint64_t n_rows = 1ULL << 27, n_rounds = 1ULL << 22;   // 2 GB table, 537 M gathers per launch
    uint8_t* tab; uint32_t* sink;
    cudaMalloc(&tab, n_rows * 16); cudaMalloc(&sink, 4);
    cudaMemset(tab, 0xA5, n_rows * 16);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);

    CUtensorMap tmap{};
    if (mode >= 2) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
        cuuint64_t gdim[2] = {4, n_rows};           // 4 x u32 = one 16-byte row
        cuuint64_t gstr[1] = {16};
        cuuint32_t box[2] = {4, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = ((EncodeTiled)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, tab, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    }
    const size_t smem = (size_t)NW * STAGES * 32 * SLOT;
    cudaFuncSetAttribute(k_gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_gather<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char* names[] = {"lsu ld.global.nc.v4", "cp.async.bulk 16B ring", "tensor tile::gather4 ring", "hybrid lsu + gather4"};
    for (uint64_t mb : {16ULL, 2048ULL}) {
        const uint64_t win_mask = (mb << 20) / 16 - 1;
        for (int occ : {2, 3, 6}) {
            if (mode && occ * smem > 200 * 1024) continue;
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (mode == 0) k_gather<0><<<sm * occ * 2, NT, 0>>>(tab, tmap, win_mask, n_rounds, sink);
                if (mode == 1) k_gather<1><<<sm * occ, NT, smem>>>(tab, tmap, win_mask, n_rounds, sink);
                if (mode == 2) k_gather<2><<<sm * occ, NT, smem>>>(tab, tmap, win_mask, n_rounds, sink);
                if (mode == 3) k_gather<3><<<sm * occ, NT, smem>>>(tab, tmap, win_mask, n_rounds, sink);
                cudaEventRecord(b);
                cudaError_t e = cudaEventSynchronize(b);
                if (e != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(e)); return 2; }
                cudaEventElapsedTime(&ms, a, b);
            }
            printf("window=%5llu MB  %-26s box_rows=%d ctas/sm=%d  %8.2f ms  %7.2f Ggather/s\n", (unsigned long long)mb, names[mode], box_rows,
                   mode ? occ : occ * 2, ms, n_rounds * 128 / ms / 1e6);
        }
    }
    printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
