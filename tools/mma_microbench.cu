// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16, no-swizzle K-major operands) as a function of
// N, M, the number of independent accumulators and the LBO/SBO pattern.  Build: nvcc -arch=sm_100a ...
#include <cstdio>
#include "../scene_3dreconstruction_mvsnet_b200/csrc/tc_common.cuh"
using namespace mvs;

__global__ void __launch_bounds__(128, 1) bench(int M, int N, int nacc, int iters, int a_stride, long long *out, int layout, int sbo, int lbo, int tf32) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x3c003c00u;  // small bf16 values
    const uint32_t b = ptx::smem_u32(&bar);
    if (threadIdx.x == 0) { ptx::mbar_init(b, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&tslot), 512);
    ptx::fence_proxy_async_smem();
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tm = tslot;
    if (threadIdx.x < 32) {   // whole warp runs the loop with warp-uniform values; one elected lane issues
        const uint32_t fmt = tf32 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t sbase = ptx::smem_u32(smem);
        const uint64_t hi = (((uint64_t)(((uint32_t)sbo >> 4) | (1u << 14))) << 32) | ((uint64_t)layout << 61);
        const uint32_t a_lo0 = (sbase >> 4) | (((uint32_t)lbo >> 4) << 16);
        const uint32_t b_lo = ((sbase + 131072) >> 4) | (((layout ? (uint32_t)lbo : (uint32_t)N * 16u) >> 4) << 16);
        const uint32_t amask = nacc - 1;   // nacc is a power of two
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t a_lo = a_lo0 + (uint32_t)(((i + u) & 15) * (a_stride >> 4));
                const uint32_t d = tm + ((i + u) & amask) * N;
                if (threadIdx.x == 0) ptx::mma_bf16_ss(d, hi | a_lo, hi | b_lo, idesc, (i + u) >= nacc);
            }
        }
        if (threadIdx.x == 0) ptx::tcgen05_commit(b);
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
    const int iters = 2048;
    struct Cfg { const char *name; int layout, sbo, lbo, tf32; };
    Cfg cfgs[] = {{"none", 0, 128, 12288, 0}, {"sw128", 2, 1024, 16, 0}, {"sw64", 4, 512, 16, 0}, {"sw32", 6, 256, 16, 0},
                  {"none-tf32", 0, 128, 12288, 1}};
    for (auto c : cfgs)
        for (int M : {128, 64})
            for (int N : {8, 16, 64, 256}) {
                if (M == 128 && N < 16) continue;
                bench<<<148, 128, 200 * 1024>>>(M, N, 1, iters, 2048, d, c.layout, c.sbo, c.lbo, c.tf32);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("%10s M=%3d N=%3d  cyc/mma %8.1f\n", c.name, M, N, (double)h[1] / iters);
            }
    return 0;
}
