// Tight-issue micro-benchmark: descriptors are kernel parameters (uniform registers), 8 tcgen05.mma per loop
// iteration issued back to back by one elected lane.
#include <cstdio>
#include "../scene_3dreconstruction_mvsnet_b200/csrc/tc_common.cuh"
using namespace mvs;

__global__ void __launch_bounds__(128, 1) bench(uint32_t idesc, uint32_t a_lo_off, uint32_t b_lo_off, uint32_t hi32, int iters,
                                                 int dstep, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x3c003c00u;
    const uint32_t b = ptx::smem_u32(&bar);
    if (threadIdx.x == 0) { ptx::mbar_init(b, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&tslot), 512);
    ptx::fence_proxy_async_smem();
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tm = __shfl_sync(0xffffffffu, tslot, 0);
    if (threadIdx.x < 32) {
        const uint32_t sb16 = ptx::smem_u32(smem) >> 4;
        const uint64_t hi = (uint64_t)hi32 << 32;
        const uint64_t ad = hi | (uint64_t)(sb16 + a_lo_off);
        const uint64_t bd = hi | (uint64_t)(sb16 + b_lo_off);
        const bool leader = ptx::elect_one();
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
            if (leader) {
#pragma unroll
                for (int u = 0; u < 8; ++u) ptx::mma_bf16_ss(tm + (u & 3) * dstep, ad + u * 128, bd, idesc, 1);
            }
        }
        if (leader) ptx::tcgen05_commit(b);
        long long t1 = clock64();
        ptx::mbar_wait(b, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4096;
    // alignment sweep (no-swizzle K-major layout, M = 128): start address of the A tile and the byte distance between its two
    // K chunks (LBO) in 16-byte units off a 128-byte boundary
    for (int N : {16, 48, 96})
        for (uint32_t lbo : {12288u, 12288u + 16u, 1632u, 1632u + 16u})
            for (uint32_t aoff : {0u, 1u, 2u, 4u}) {
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                const uint32_t hi32 = (128u >> 4) | (1u << 14);
                const uint32_t a_lo = ((lbo >> 4) << 16) + aoff, b_lo = (131072u >> 4) | ((((uint32_t)N * 16) >> 4) << 16);
                bench<<<148, 128, 200 * 1024>>>(idesc, a_lo, b_lo, hi32, iters, N, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("N=%3d lbo=%5u aoff=%u*16B  total %7.1f cyc/mma\n", N, lbo, aoff, (double)h[1] / iters);
            }
    return 0;
}
