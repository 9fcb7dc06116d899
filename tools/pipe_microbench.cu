// Issue / pipe rates of the instruction kinds the fused warp+variance kernel is made of, on sm_100a.
// One CTA of 512 threads per SM (4 warps per scheduler), every thread runs ITER iterations of a body of UNROLL
// independent dependency chains; the result is reported as warp-instructions per clock per SM (clock64 inside
// the kernel, slowest CTA), so it does not depend on the SM frequency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_microbench tools/pipe_microbench.cu && tools/pipe_microbench
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int kThreads = 512;
constexpr int kIter = 4096;

enum Mix {
    HFMA2_ONLY, FFMA_ONLY, HFMA2_FFMA, HFMA2_FADD, HFMA2_LOP3, F2FP_ONLY, HFMA2_F2FP, FMNMX_ONLY, HFMA2_FMNMX, HMNMX2_ONLY,
    FRND_ONLY, F2I_ONLY, MUFU_ONLY, IMAD_ONLY, HFMA2_IMAD, SHFL_ONLY, LDS128_ONLY, LDS128_HALFPRED_Q, LDS128_HALFPRED_I,
    HFMA2_LDS128_6to1, HMMA_F16ACC, HMMA_F32ACC, HFMA2_HMMA, FHFMA_ONLY, HADD2_ONLY, HFMA2_HADD2, LDSM_X4, FFMA2_ONLY, HFMA2_RELU,
    LEA_ONLY, NUM_MIX
};
const char *kNames[NUM_MIX] = {
    "HFMA2", "FFMA (3 regs)", "HFMA2 + FFMA 1:1", "HFMA2 + FADD 1:1", "HFMA2 + LOP3 1:1", "F2FP.F16.PACK", "HFMA2 + F2FP 1:1",
    "FMNMX", "HFMA2 + FMNMX 1:1", "HMNMX2", "FRND.FLOOR", "F2I.FLOOR", "MUFU.RCP", "IMAD", "HFMA2 + IMAD 1:1", "SHFL.BFLY",
    "LDS.128", "LDS.128 lanes 0-15 only", "LDS.128 even lanes only", "HFMA2 + LDS.128 6:1", "HMMA.16816.F16", "HMMA.16816.F32",
    "HFMA2 + HMMA.F16 4:1", "FHFMA (fma.f32.f16)", "HADD2", "HFMA2 + HADD2 1:1", "LDSM.x4.trans", "FFMA2", "HFMA2.RELU",
    "LEA"};

__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int MIX>
__global__ void __launch_bounds__(kThreads, 1) bench(float *out, long long *cyc, int iters) {
    __shared__ __align__(128) uint4 tile[2048];
    for (int i = threadIdx.x; i < 2048; i += kThreads) tile[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t h[8];
    float f[8];
    int n[8];
    uint32_t mm[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        h[i] = 0x3c003c00u + threadIdx.x + i;
        f[i] = 1.0f + 1e-3f * (threadIdx.x + i);
        n[i] = threadIdx.x * 3 + i;
        for (int j = 0; j < 4; ++j) mm[i][j] = 0x3c003800u + i + j;
    }
    const uint32_t hb = 0x3bff3bffu, hc = 0x10001000u;
    const float fb = 0.99991f, fc = 1e-4f;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(tile) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512;
    uint4 acc4 = make_uint4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MIX == HFMA2_ONLY || MIX == HFMA2_FFMA || MIX == HFMA2_FADD || MIX == HFMA2_LOP3 || MIX == HFMA2_F2FP ||
                MIX == HFMA2_FMNMX || MIX == HFMA2_IMAD || MIX == HFMA2_HADD2)
                h[i] = hfma2(h[i], hb, hc);
            if (MIX == FFMA_ONLY || MIX == HFMA2_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(fc));
            if (MIX == HFMA2_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fc));
            if (MIX == HFMA2_LOP3) asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[i]) : "r"(lane + 77));
            if (MIX == F2FP_ONLY || MIX == HFMA2_F2FP) {
                uint32_t r;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(f[i]), "f"(f[(i + 1) & 7]));
                n[i] ^= r;
            }
            if (MIX == FMNMX_ONLY || MIX == HFMA2_FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fc * (float)(it & 3)));
            if (MIX == HMNMX2_ONLY) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hc + it));
            if (MIX == FRND_ONLY) asm volatile("cvt.rmi.f32.f32 %0, %0;" : "+f"(f[i]));
            if (MIX == F2I_ONLY) {
                int r;
                asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(r) : "f"(f[i]));
                n[i] += r;
            }
            if (MIX == MUFU_ONLY) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            if (MIX == IMAD_ONLY || MIX == HFMA2_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(n[i]) : "r"(lane | 3), "r"(it));
            if (MIX == SHFL_ONLY) n[i] = __shfl_xor_sync(0xffffffffu, n[i], 1 + (i & 3));
            if (MIX == FHFMA_ONLY) asm volatile("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(f[i]) : "h"((unsigned short)(0x3800 + i)));
            if (MIX == HADD2_ONLY || MIX == HFMA2_HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[(i + 4) & 7]) : "r"(hc));
            if (MIX == FFMA2_ONLY) {
                unsigned long long v = ((unsigned long long)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i]);
                unsigned long long b2 = ((unsigned long long)__float_as_uint(fb) << 32) | __float_as_uint(fb);
                unsigned long long c2 = ((unsigned long long)__float_as_uint(fc) << 32) | __float_as_uint(fc);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(b2), "l"(c2));
                f[i] = __uint_as_float((uint32_t)v);
            }
            if (MIX == HFMA2_RELU) asm volatile("fma.rn.relu.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(hb), "r"(hc));
            if (MIX == LEA_ONLY) asm volatile("{ .reg .b32 t; shl.b32 t, %0, 4; add.s32 %0, t, %1; }" : "+r"(n[i]) : "r"(lane));
            if (MIX == LDS128_ONLY || MIX == LDS128_HALFPRED_Q || MIX == LDS128_HALFPRED_I) {
                const bool on = MIX == LDS128_ONLY || (MIX == LDS128_HALFPRED_Q ? lane < 16 : (lane & 1) == 0);
                if (on) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sbase + ((i * 528 + it * 16) & 0x3ff0)));
                    acc4.x ^= v.x; acc4.y ^= v.y; acc4.z ^= v.z; acc4.w ^= v.w;
                }
            }
            if (MIX == LDSM_X4) {
                uint32_t a, b, c, d;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(sbase + ((i * 528 + it * 16) & 0x3ff0)));
                acc4.x ^= a; acc4.y ^= b; acc4.z ^= c; acc4.w ^= d;
            }
            if (MIX == HMMA_F16ACC || MIX == HFMA2_HMMA) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                             : "+r"(mm[i][0]), "+r"(mm[i][1]) : "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(hb), "r"(hc));
            }
            if (MIX == HMMA_F32ACC) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i & 3]), "+f"(f[(i & 3) + 4]), "+f"(*(float *)&mm[i][2]), "+f"(*(float *)&mm[i][3])
                             : "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(hb), "r"(hc));
            }
        }
        if (MIX == HFMA2_LDS128_6to1) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int k = 0; k < 6; ++k) h[i] = hfma2(h[i], hb, hc);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sbase + ((i * 528 + it * 16) & 0x3ff0)));
                acc4.x ^= v.x; acc4.y ^= v.y; acc4.z ^= v.z; acc4.w ^= v.w;
            }
        }
        if (MIX == HFMA2_HMMA) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) h[i] = hfma2(h[i], hb, hc);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i] + (float)n[i] + (float)h[i] + (float)(mm[i][0] ^ mm[i][1] ^ mm[i][2] ^ mm[i][3]);
    out[blockIdx.x * kThreads + threadIdx.x] = s + (float)(acc4.x ^ acc4.y ^ acc4.z ^ acc4.w);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MIX>
void run(float *out, long long *cyc, int nsm) {
    // instructions per thread per iteration (of the kinds named, ignoring loop overhead)
    double per_iter = 8;
    if (MIX == HFMA2_FFMA || MIX == HFMA2_FADD || MIX == HFMA2_LOP3 || MIX == HFMA2_F2FP || MIX == HFMA2_FMNMX ||
        MIX == HFMA2_IMAD || MIX == HFMA2_HADD2) per_iter = 16;
    if (MIX == HFMA2_LDS128_6to1) per_iter = 56;
    if (MIX == HFMA2_HMMA) per_iter = 40;
    if (MIX == LEA_ONLY) per_iter = 16;
    bench<MIX><<<nsm, kThreads>>>(out, cyc, 64);
    bench<MIX><<<nsm, kThreads>>>(out, cyc, kIter);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s CUDA error %s\n", kNames[MIX], cudaGetErrorString(e)); return; }
    static long long hc[1024];
    cudaMemcpy(hc, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < nsm; ++i) mx = hc[i] > mx ? hc[i] : mx;
    const double winst = per_iter * kIter * (kThreads / 32);
    printf("%-28s %8.3f warp-inst/clk/SM   (%lld cycles)\n", kNames[MIX], winst / (double)mx, mx);
}

template <int M>
struct RunAll {
    static void go(float *out, long long *cyc, int nsm) {
        run<M>(out, cyc, nsm);
        RunAll<M + 1>::go(out, cyc, nsm);
    }
};
template <>
struct RunAll<NUM_MIX> {
    static void go(float *, long long *, int) {}
};

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    float *out;
    long long *cyc;
    cudaMalloc(&out, (size_t)nsm * kThreads * 4);
    cudaMalloc(&cyc, nsm * sizeof(long long));
    printf("%s, %d SMs; one CTA of %d threads per SM; LDS rows: 128 B/clk/SM = 0.25 LDS.128 warp-inst/clk\n", p.name, nsm, kThreads);
    RunAll<0>::go(out, cyc, nsm);
    return 0;
}
