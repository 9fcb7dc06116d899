// CostRegNet building blocks, fp32 CUDA-core path (bit-for-tolerance parity with the reference).
//
// Replaces, for the reference:
//   models/module.py:26-33   ConvBnReLU3D  = nn.Conv3d(k3,p1,s1|2,bias=False) + BatchNorm3d + ReLU
//   models/mvsnet.py:46-59   nn.ConvTranspose3d(k3,s2,p1,op1,bias=False) + BatchNorm3d + ReLU
//   models/mvsnet.py:62      prob = nn.Conv3d(8,1,3,padding=1)  (bias, no BN/ReLU)
//   models/mvsnet.py:69-71   skip additions  conv4 + conv7(x) ...
// Eval-mode BatchNorm is folded into the weights/shift by the caller, so BN, ReLU and the skip add
// cost no extra pass over the activation (the reference runs each as a separate kernel).
//
// This file is the strict-precision (fp32 FMA) path.  The tensor-core path is conv3d_tc.cu.
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mvs {

// 4-byte asynchronous copy global -> shared; src_bytes = 0 writes a zero instead (nothing is read)
__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// Direct 3x3x3 convolution, NCDHW.  One CTA: TZ x TY x 32 output voxels x COUT_T output channels.
// Thread: 4 consecutive x outputs x COUT_T channels in registers.  Input channels are processed in
// chunks of CK through a shared-memory halo tile; weights of the chunk are broadcast from smem.
// ------------------------------------------------------------------------------------------------
template <int S, int TZ, int TY>
struct ConvTile {
    static constexpr int IZ = (TZ - 1) * S + 3;
    static constexpr int IY = (TY - 1) * S + 3;
    static constexpr int IX = 31 * S + 3;
    static constexpr int IXP = (IX + 3) / 4 * 4;
    static constexpr int PER_CH = IZ * IY * IXP;
    static constexpr int THREADS = TZ * TY * 8;
};

// shift (+ ReLU) and the 4-wide store of a thread's outputs
template <int COUT_T, bool SCALAR = false>
__device__ __forceinline__ void conv_store(const float2 (&acc)[4][(COUT_T + 1) / 2], const float *__restrict__ shift, int relu,
                                           float *__restrict__ y, int b, int co0, int Cout, int oz, int oy, int ox, int Do, int Ho,
                                           int Wo) {
    if (oz >= Do || oy >= Ho || ox >= Wo) return;   // (ox may be negative in the SCALAR form: per-element test below)
    const size_t out_cs = (size_t)Do * Ho * Wo;
#pragma unroll
    for (int q = 0; q < COUT_T; ++q) {
        if (co0 + q >= Cout) break;
        const float sh = __ldg(shift + co0 + q);
        float v[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            v[o] = ((q & 1) ? acc[o][q >> 1].y : acc[o][q >> 1].x) + sh;
            if (relu) v[o] = fmaxf(v[o], 0.f);
        }
        float *op = y + ((size_t)b * Cout + co0 + q) * out_cs + ((size_t)oz * Ho + oy) * Wo + ox;
        if (!SCALAR && (Wo & 3) == 0) {
            *reinterpret_cast<float4 *>(op) = make_float4(v[0], v[1], v[2], v[3]);
        } else if (SCALAR) {
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (ox + o >= 0 && ox + o < Wo) op[o] = v[o];
        } else {
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (ox + o < Wo) op[o] = v[o];
        }
    }
}

// One chunk of CK input channels from a halo tile in shared memory: the 27 taps of every channel, 4 x COUT_T FMAs per tap and
// thread, in the order (channel, kd, kh, kw) -- both kernels below accumulate through this function, so their results are
// bit-identical.
// XO: column of the tile row at which a thread's first input sits, relative to tx * 4 * S (0: the tile starts at the first
// needed column; 3 (stride 2, TMA tiles): the tile starts 3 columns earlier, see conv3d_fp32_tma_kernel).
template <int S, int CK, int TZ, int TY, int COUT_T, int XO = 0>
__device__ __forceinline__ void conv_chunk(const float *__restrict__ s_in, const float *__restrict__ s_w, int tx, int ty, int tz,
                                           float2 (&acc)[4][(COUT_T + 1) / 2]) {
    using T = ConvTile<S, TZ, TY>;
#pragma unroll 1
    for (int c = 0; c < CK; ++c) {
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float *row = s_in + ((c * T::IZ + tz * S + kd) * T::IY + ty * S + kh) * T::IXP + tx * 4 * S;
                float in[4 * S + 4];
                if (S == 1) {
                    const float4 a = *reinterpret_cast<const float4 *>(row);
                    const float2 e = *reinterpret_cast<const float2 *>(row + 4);
                    in[0] = a.x; in[1] = a.y; in[2] = a.z; in[3] = a.w; in[4] = e.x; in[5] = e.y;
                } else if (XO == 0) {
                    const float4 a = *reinterpret_cast<const float4 *>(row);
                    const float4 e = *reinterpret_cast<const float4 *>(row + 4);
                    in[0] = a.x; in[1] = a.y; in[2] = a.z; in[3] = a.w;
                    in[4] = e.x; in[5] = e.y; in[6] = e.z; in[7] = e.w;
                    in[8] = row[8];
                } else {
                    static_assert(XO == 0 || XO == 3, "tile column offset");
                    const float4 a = *reinterpret_cast<const float4 *>(row + 4);
                    const float4 e = *reinterpret_cast<const float4 *>(row + 8);
                    in[0] = row[3];
                    in[1] = a.x; in[2] = a.y; in[3] = a.z; in[4] = a.w;
                    in[5] = e.x; in[6] = e.y; in[7] = e.z; in[8] = e.w;
                }
                const float *wp = s_w + (c * 27 + (kd * 3 + kh) * 3) * COUT_T;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    float wr[COUT_T];
                    if (COUT_T % 4 == 0) {
#pragma unroll
                        for (int q = 0; q < COUT_T / 4; ++q) {
                            const float4 t = *reinterpret_cast<const float4 *>(wp + kw * COUT_T + 4 * q);
                            wr[4 * q] = t.x; wr[4 * q + 1] = t.y; wr[4 * q + 2] = t.z; wr[4 * q + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < COUT_T; ++q) wr[q] = wp[kw * COUT_T + q];
                    }
                    if constexpr (COUT_T % 2 == 0) {
                        // packed fp32 pairs (FFMA2): the same two IEEE fmas per instruction, half the issue slots -- the
                        // scalar form of this loop ran at 82 % issue-slot utilisation with the FMA pipe at 71 % (ncu, conv0)
#pragma unroll
                        for (int o = 0; o < 4; ++o) {
                            const float2 a2 = make_float2(in[o * S + kw], in[o * S + kw]);
#pragma unroll
                            for (int q = 0; q < COUT_T / 2; ++q)
                                acc[o][q] = __ffma2_rn(a2, make_float2(wr[2 * q], wr[2 * q + 1]), acc[o][q]);
                        }
                    } else {
#pragma unroll
                        for (int o = 0; o < 4; ++o)
#pragma unroll
                            for (int q = 0; q < COUT_T; ++q) {
                                float &d = (q & 1) ? acc[o][q >> 1].y : acc[o][q >> 1].x;
                                d = fmaf(in[o * S + kw], wr[q], d);
                            }
                    }
                }
            }
        }
    }
}

template <int S, int CK, int TZ, int TY, int COUT_T>
__global__ void __launch_bounds__(TZ *TY * 8)
conv3d_fp32_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ shift, int relu,
                   float *__restrict__ y, int Cin, int Cout, int Din, int Hin, int Win, int Do, int Ho, int Wo,
                   int tiles_x) {
    using T = ConvTile<S, TZ, TY>;
    extern __shared__ __align__(128) float smem[];
    float *s_in = smem;                      // [CK][IZ][IY][IXP]
    float *s_w = smem + CK * T::PER_CH;      // [CK][27][COUT_T]

    const int tid = threadIdx.x;
    const int tx = tid & 7, ty = (tid >> 3) % TY, tz = tid / (8 * TY);
    const int cgroups = (Cout + COUT_T - 1) / COUT_T;
    const int b = blockIdx.z / cgroups;
    const int co0 = (blockIdx.z % cgroups) * COUT_T;
    const int ox0 = (blockIdx.x % tiles_x) * 32, oy0 = (blockIdx.x / tiles_x) * TY, oz0 = blockIdx.y * TZ;
    const int ix0 = ox0 * S - 1, iy0 = oy0 * S - 1, iz0 = oz0 * S - 1;
    const size_t in_cs = (size_t)Din * Hin * Win;

    float2 acc[4][(COUT_T + 1) / 2];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int c = 0; c < (COUT_T + 1) / 2; ++c) acc[o][c] = make_float2(0.f, 0.f);

    for (int ci0 = 0; ci0 < Cin; ci0 += CK) {
        __syncthreads();
        // halo tile, one row of IX floats per warp iteration: the (channel, z, y) decode and the y / z bounds are
        // warp-uniform, a lane only adds its x, and the copies are asynchronous (cp.async with zero fill outside the volume), so a
        // whole tile is in flight before anything waits.  (An element-indexed loop of __ldg + st.shared spent ~60 instructions per element on
        // div / mod chains and 64-bit address arithmetic -- a third of this kernel's instructions at stride 1 and
        // three quarters at stride 2, where the halo is 4x the outputs.)
        {
            constexpr int NW = T::THREADS / 32;
            const int wid = tid >> 5, ln = tid & 31;
            for (int row = wid; row < CK * T::IZ * T::IY; row += NW) {
                const int yy = row % T::IY;
                const int r2 = row / T::IY;
                const int zz = r2 % T::IZ;
                const int c = r2 / T::IZ;
                const int gy = iy0 + yy, gz = iz0 + zz;
                const bool row_ok = (ci0 + c < Cin) && gy >= 0 && gy < Hin && gz >= 0 && gz < Din;
                const float *src = x + ((size_t)b * Cin + ci0 + c) * in_cs + ((size_t)(row_ok ? gz : 0) * Hin + (row_ok ? gy : 0)) * Win;
                float *dst = s_in + ((c * T::IZ + zz) * T::IY + yy) * T::IXP;
#pragma unroll
                for (int xx = ln; xx < T::IX; xx += 32) {
                    const int gx = ix0 + xx;
                    const bool ok = row_ok && gx >= 0 && gx < Win;
                    cp_async4(dst + xx, src + (ok ? gx : 0), ok ? 4 : 0);   // asynchronous: every row of the tile in flight at once
                }
            }
        }
        for (int idx = tid; idx < CK * 27 * COUT_T; idx += T::THREADS) {
            const int co = idx % COUT_T;
            const int tap = (idx / COUT_T) % 27;
            const int c = idx / (COUT_T * 27);
            float v = 0.f;
            if (co0 + co < Cout && ci0 + c < Cin) v = __ldg(w + ((size_t)(co0 + co) * Cin + ci0 + c) * 27 + tap);
            s_w[idx] = v;
        }
        cp_async_wait_all();
        __syncthreads();

        conv_chunk<S, CK, TZ, TY, COUT_T>(s_in, s_w, tx, ty, tz, acc);
    }

    conv_store<COUT_T>(acc, shift, relu, y, b, co0, Cout, oz0 + tz, oy0 + ty, ox0 + tx * 4, Do, Ho, Wo);
}

// ------------------------------------------------------------------------------------------------
// The same tile and arithmetic with the halo tile brought by TMA: one 5-D box (x, y, z, channel chunk, batch element) per
// chunk, zero-filled outside the volume by the copy engine (= the convolution's zero padding), two stages so that the next
// chunk is in flight while this one is computed.  The element-wise cp.async loop of the kernel above executes ~190
// instructions per tile row (bounds, 64-bit addresses, one 4-byte copy per lane): 43 % of all instructions of conv0 at the DTU
// shape (ncu, profiles/r02z).  Needs W % 4 == 0 and a 16-byte aligned input (tensor-map strides); other shapes take the kernel
// above.  Weights of the next chunk arrive by cp.async next to the tile.
// ------------------------------------------------------------------------------------------------
template <int S, int CK, int TZ, int TY, int COUT_T>
__global__ void __launch_bounds__(TZ *TY * 8)
conv3d_fp32_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ w, const float *__restrict__ shift,
                       int relu, float *__restrict__ y, int Cin, int Cout, int Do, int Ho, int Wo, int tiles_x) {
    using T = ConvTile<S, TZ, TY>;
    // The innermost start coordinate of a TMA box must be a multiple of 16 bytes: the box starts at column S * ox0 - 4, not
    // S * ox0 - 1.  Stride 1: the OUTPUT tile moves 3 columns to the left with it (ox0 - 3 ...), so a thread's six inputs stay
    // an aligned float4 + float2 of its row and only its four stores become scalar.  Stride 2: inputs sit at column 3 of the
    // thread's span (scalar + 2 x float4 instead of 2 x float4 + scalar).
    constexpr int kOutShift = (S == 1) ? 3 : 0;
    constexpr int kXO = (S == 1) ? 0 : 3;
    constexpr int kBox = CK * T::PER_CH;            // floats a TMA box delivers
    constexpr int kTile = (kBox + 31) / 32 * 32;    // floats per stage: 128-byte aligned TMA destinations
    constexpr int kW = CK * 27 * COUT_T;
    extern __shared__ __align__(128) float smem[];
    // TMA destinations must be 128-byte aligned: align by hand (the launch allocates 128 bytes of slack)
    float *s_in = smem + ((128u - (ptx::smem_u32(smem) & 127u)) & 127u) / 4;   // [2][CK][IZ][IY][IXP]
    float *s_w = s_in + 2 * kTile;                                            // [2][CK][27][COUT_T]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_w + 2 * kW);             // [2]  (2 * kW floats: a multiple of 8 bytes)

    const int tid = threadIdx.x;
    const int tx = tid & 7, ty = (tid >> 3) % TY, tz = tid / (8 * TY);
    const int cgroups = (Cout + COUT_T - 1) / COUT_T;
    const int b = blockIdx.z / cgroups;
    const int co0 = (blockIdx.z % cgroups) * COUT_T;
    const int ox0 = (blockIdx.x % tiles_x) * 32, oy0 = (blockIdx.x / tiles_x) * TY, oz0 = blockIdx.y * TZ;
    const int nchunks = (Cin + CK - 1) / CK;
    const uint32_t bar0 = ptx::smem_u32(s_bar), in0 = ptx::smem_u32(s_in);

    if (tid == 0) {
        // a stage is complete when the TMA box has landed (one arrival + its bytes) and every thread's weight copies have
        // (cp.async.mbarrier.arrive.noinc: one arrival per thread when its earlier cp.async are done)
        ptx::mbar_init(bar0, 1 + T::THREADS);
        ptx::mbar_init(bar0 + 8, 1 + T::THREADS);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&tmap);
    }
    __syncthreads();

    const CUtensorMap *const tm = &tmap;  // (taken here, not inside a lambda: the descriptor must stay in parameter space)
    auto issue = [=](int k) {
        const int st = k & 1, ci0 = k * CK;
        if (tid == 0) {
            ptx::fence_proxy_async_smem();  // the stage was read through the generic proxy two chunks ago
            ptx::mbar_arrive_expect_tx(bar0 + 8 * st, (uint32_t)(kBox * 4));
            ptx::tma_load_5d(in0 + (uint32_t)(st * kTile * 4), tm, bar0 + 8 * st, ox0 * S - 4, oy0 * S - 1, oz0 * S - 1, ci0, b);
        }
        float *dw = s_w + st * kW;
        for (int idx = tid; idx < kW; idx += T::THREADS) {
            const int co = idx % COUT_T;
            const int tap = (idx / COUT_T) % 27;
            const int c = idx / (COUT_T * 27);
            const bool ok = co0 + co < Cout && ci0 + c < Cin;
            cp_async4(dw + idx, w + (ok ? ((size_t)(co0 + co) * Cin + ci0 + c) * 27 + tap : 0), ok ? 4 : 0);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar0 + 8 * st) : "memory");
    };

    float2 acc[4][(COUT_T + 1) / 2];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int c = 0; c < (COUT_T + 1) / 2; ++c) acc[o][c] = make_float2(0.f, 0.f);

    issue(0);
    for (int k = 0; k < nchunks; ++k) {
        const int st = k & 1;
        if (k + 1 < nchunks) issue(k + 1);  // into the other stage: its readers finished behind the barrier that ended chunk k - 1
        ptx::mbar_wait(bar0 + 8 * st, (uint32_t)((k >> 1) & 1));  // tile and weights of this chunk are in shared memory
        conv_chunk<S, CK, TZ, TY, COUT_T, kXO>(s_in + st * kTile, s_w + st * kW, tx, ty, tz, acc);
        __syncthreads();  // the stage may be refilled
    }
    conv_store<COUT_T, S == 1>(acc, shift, relu, y, b, co0, Cout, oz0 + tz, oy0 + ty, ox0 - kOutShift + tx * 4, Do, Ho, Wo);
}

template <int S, int CK, int TZ, int TY, int COUT_T>
static int launch_conv(const float *x, const float *w, const float *shift, int relu, float *y, int B, int Cin, int Cout,
                       int D, int H, int W, cudaStream_t st) {
    using T = ConvTile<S, TZ, TY>;
    const int Do = (D - 1) / S + 1, Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1;
    const int tiles_x = cdiv(Wo, 32), tiles_y = cdiv(Ho, TY), tiles_z = cdiv(Do, TZ);
    const int cgroups = cdiv(Cout, COUT_T);
    MVS_REQUIRE(tiles_z <= 65535 && (long long)B * cgroups <= 65535, "conv3d: grid too large");
    const dim3 grid(tiles_x * tiles_y, tiles_z, B * cgroups);
    if ((W & 3) == 0 && ((uintptr_t)x & 15) == 0) {
        // halo tiles by TMA (tensor-map strides must be multiples of 16 bytes)
        tmap_encode_fn enc = get_tmap_encode();
        MVS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
        CUtensorMap tmap;
        const cuuint64_t gdim[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)Cin, (cuuint64_t)B};
        const cuuint64_t gstr[4] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4,
                                    (cuuint64_t)Cin * D * H * W * 4};
        const cuuint32_t box[5] = {(cuuint32_t)T::IXP, (cuuint32_t)T::IY, (cuuint32_t)T::IZ, (cuuint32_t)CK, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float *>(x), gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return set_error(MVS_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 conv input) failed (%d)", (int)cr);
        const size_t smem = (size_t)2 * ((CK * T::PER_CH + 31) / 32 * 32 + CK * 27 * COUT_T) * sizeof(float) + 16 + 128;
        auto kern = conv3d_fp32_tma_kernel<S, CK, TZ, TY, COUT_T>;
        MVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int tiles_xt = (S == 1) ? cdiv(Wo + 3, 32) : tiles_x;  // stride 1: tiles start at 32 i - 3
        kern<<<dim3(tiles_xt * tiles_y, tiles_z, B * cgroups), T::THREADS, smem, st>>>(tmap, w, shift, relu, y, Cin, Cout, Do, Ho, Wo,
                                                                                     tiles_xt);
        MVS_LAUNCH_CHECK(1);
        return MVS_OK;
    }
    const size_t smem = (size_t)(CK * T::PER_CH + CK * 27 * COUT_T) * sizeof(float);
    auto kern = conv3d_fp32_kernel<S, CK, TZ, TY, COUT_T>;
    if (smem > 48 * 1024) MVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, T::THREADS, smem, st>>>(x, w, shift, relu, y, Cin, Cout, D, H, W, Do, Ho, Wo, tiles_x);
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

// ------------------------------------------------------------------------------------------------
// Transposed convolution k3 s2 p1 op1 as a gather: out[o] = sum_{i,k : o = 2i - 1 + k} in[i] w[k].
// Per dimension an even output (parity 0) sees tap k=1 of input i=o/2; an odd output (parity 1)
// sees k=2 of input (o-1)/2 and k=0 of input (o+1)/2.  One thread owns one low-resolution position
// and produces its 2x2x2 output block for COUT_T channels from the 2x2x2 input neighbourhood.
// ------------------------------------------------------------------------------------------------
template <int COUT_T, int CK>
__global__ void __launch_bounds__(128)
convT3d_fp32_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ shift, int relu,
                    const float *__restrict__ skip, float *__restrict__ y, int Cin, int Cout, int D, int H, int W) {
    __shared__ __align__(16) float s_w[CK * 27 * COUT_T];  // [ci][tap][co]
    const int tid = threadIdx.x;
    const int cgroups = (Cout + COUT_T - 1) / COUT_T;
    const int b = blockIdx.z / cgroups;
    const int co0 = (blockIdx.z % cgroups) * COUT_T;
    const int z = blockIdx.y;
    const int pos = blockIdx.x * 128 + tid;
    const bool live = pos < H * W;
    const int yy = live ? pos / W : 0, xx = live ? pos % W : 0;
    const size_t in_cs = (size_t)D * H * W;
    const bool hz = z + 1 < D, hy = yy + 1 < H, hx = xx + 1 < W;

    float acc[8][COUT_T];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int q = 0; q < COUT_T; ++q) acc[p][q] = 0.f;

    for (int ci0 = 0; ci0 < Cin; ci0 += CK) {
        __syncthreads();
        for (int idx = tid; idx < CK * 27 * COUT_T; idx += 128) {
            const int co = idx % COUT_T;
            const int tap = (idx / COUT_T) % 27;
            const int c = idx / (COUT_T * 27);
            float v = 0.f;
            if (co0 + co < Cout && ci0 + c < Cin) v = __ldg(w + ((size_t)(ci0 + c) * Cout + co0 + co) * 27 + tap);
            s_w[idx] = v;
        }
        __syncthreads();
        if (!live) continue;
#pragma unroll 1
        for (int c = 0; c < CK; ++c) {
            if (ci0 + c >= Cin) break;
            const float *ip = x + ((size_t)b * Cin + ci0 + c) * in_cs + ((size_t)z * H + yy) * W + xx;
            float in[8];  // [dz][dy][dx]
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int dz = n >> 2, dy = (n >> 1) & 1, dx = n & 1;
                const bool ok = (!dz || hz) && (!dy || hy) && (!dx || hx);
                in[n] = ok ? __ldg(ip + ((size_t)dz * H + dy) * W + dx) : 0.f;
            }
            const float *wc = s_w + c * 27 * COUT_T;
#pragma unroll
            for (int p = 0; p < 8; ++p) {           // output parity (pz,py,px)
                const int pz = p >> 2, py = (p >> 1) & 1, px = p & 1;
#pragma unroll
                for (int n = 0; n < 8; ++n) {       // input offset (dz,dy,dx)
                    const int dz = n >> 2, dy = (n >> 1) & 1, dx = n & 1;
                    if ((dz && !pz) || (dy && !py) || (dx && !px)) continue;  // even outputs only see offset 0
                    const int kd = pz ? (dz ? 0 : 2) : 1;
                    const int kh = py ? (dy ? 0 : 2) : 1;
                    const int kw = px ? (dx ? 0 : 2) : 1;
                    const float *wt = wc + ((kd * 3 + kh) * 3 + kw) * COUT_T;
#pragma unroll
                    for (int q = 0; q < COUT_T; ++q) acc[p][q] = fmaf(in[n], wt[q], acc[p][q]);
                }
            }
        }
    }
    if (!live) return;
    const int Ho = 2 * H, Wo = 2 * W;
    const size_t out_cs = (size_t)8 * in_cs;
#pragma unroll
    for (int q = 0; q < COUT_T; ++q) {
        if (co0 + q >= Cout) break;
        const float sh = __ldg(shift + co0 + q);
#pragma unroll
        for (int pzy = 0; pzy < 4; ++pzy) {
            const int pz = pzy >> 1, py = pzy & 1;
            const size_t o = ((size_t)b * Cout + co0 + q) * out_cs + ((size_t)(2 * z + pz) * Ho + 2 * yy + py) * Wo + 2 * xx;
            float v0 = acc[pzy * 2 + 0][q] + sh, v1 = acc[pzy * 2 + 1][q] + sh;
            if (relu) {
                v0 = fmaxf(v0, 0.f);
                v1 = fmaxf(v1, 0.f);
            }
            if (skip) {  // skip + relu(bn(convT))   (mvsnet.py:69-71)
                const float2 s = __ldg(reinterpret_cast<const float2 *>(skip + o));
                v0 += s.x;
                v1 += s.y;
            }
            *reinterpret_cast<float2 *>(y + o) = make_float2(v0, v1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same gather with two x-adjacent low-resolution positions and four output channels per thread.  The one-position form
// above re-reads all 27 weight vectors per input channel for 216 FMAs -- one broadcast LDS.128 per four FMAs, which is the
// shared-memory pipe's whole budget (ncu: 20 TFLOP/s).  Two positions share every weight vector (one LDS.128 per eight FMAs)
// and 4 of their 12 inputs; the FMAs are packed pairs of output channels (FFMA2).  Accumulation order per output element is
// that of the kernel above (input channel, then input offset): bit-identical results.
// ------------------------------------------------------------------------------------------------
template <int CK>
__global__ void __launch_bounds__(128)
convT3d_fp32_pair_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ shift, int relu,
                         const float *__restrict__ skip, float *__restrict__ y, int Cin, int Cout, int D, int H, int W) {
    constexpr int CT = 4;
    __shared__ __align__(16) float s_w[CK * 27 * CT];  // [ci][tap][co]
    // the thread's 16 skip float4 (4 channels x 4 (pz, py) rows), fetched asynchronously before the accumulation and read back
    // in the epilogue: with 12 warps per SM nothing else hides a DRAM round trip in front of the stores
    __shared__ __align__(16) float4 s_skip[16][128];
    const int tid = threadIdx.x;
    const int cgroups = (Cout + CT - 1) / CT;
    const int b = blockIdx.z / cgroups;
    const int co0 = (blockIdx.z % cgroups) * CT;
    const int z = blockIdx.y;
    const int Wp = (W + 1) >> 1;  // position pairs per row
    const int pos = blockIdx.x * 128 + tid;
    const bool live = pos < H * Wp;
    const int yy = live ? pos / Wp : 0, xx = live ? 2 * (pos % Wp) : 0;
    const size_t in_cs = (size_t)D * H * W;
    const bool hz = z + 1 < D, hy = yy + 1 < H;
    const bool hx1 = xx + 1 < W, hx2 = xx + 2 < W;

    const int Ho = 2 * H, Wo = 2 * W;
    const size_t out_cs = (size_t)8 * in_cs;
    const bool vec4 = hx1 && (W & 1) == 0;  // four consecutive outputs, 16-byte aligned (xx even, Wo a multiple of 4)
    const bool skip_staged = skip != nullptr && live && vec4;
    if (skip_staged) {
#pragma unroll
        for (int q = 0; q < CT; ++q)
#pragma unroll
            for (int pzy = 0; pzy < 4; ++pzy) {
                const size_t o = ((size_t)b * Cout + min(co0 + q, Cout - 1)) * out_cs +
                                 ((size_t)(2 * z + (pzy >> 1)) * Ho + 2 * yy + (pzy & 1)) * Wo + 2 * xx;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_skip[q * 4 + pzy][tid])),
                             "l"(skip + o)
                             : "memory");
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    float2 acc[2][8][CT / 2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int q = 0; q < CT / 2; ++q) acc[j][p][q] = make_float2(0.f, 0.f);

    for (int ci0 = 0; ci0 < Cin; ci0 += CK) {
        __syncthreads();
        for (int idx = tid; idx < CK * 27 * CT; idx += 128) {
            const int co = idx % CT;
            const int tap = (idx / CT) % 27;
            const int c = idx / (CT * 27);
            float v = 0.f;
            if (co0 + co < Cout && ci0 + c < Cin) v = __ldg(w + ((size_t)(ci0 + c) * Cout + co0 + co) * 27 + tap);
            s_w[idx] = v;
        }
        __syncthreads();
        if (!live) continue;
        // the 12 inputs of a channel are loaded one channel ahead: with 3 CTAs of 4 warps per SM nothing else hides their latency
        auto load_inputs = [&](int ci, float (&v)[2][2][3]) {
            const float *ip = x + ((size_t)b * Cin + ci) * in_cs + ((size_t)z * H + yy) * W + xx;
#pragma unroll
            for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                for (int dy = 0; dy < 2; ++dy) {
                    const bool okr = (ci < Cin) && (!dz || hz) && (!dy || hy);
                    const float *r = ip + ((size_t)dz * H + dy) * W;
                    v[dz][dy][0] = okr ? __ldg(r) : 0.f;
                    v[dz][dy][1] = (okr && hx1) ? __ldg(r + 1) : 0.f;
                    v[dz][dy][2] = (okr && hx2) ? __ldg(r + 2) : 0.f;
                }
        };
        float nxt[2][2][3];
        load_inputs(ci0, nxt);
#pragma unroll 1
        for (int c = 0; c < CK; ++c) {
            if (ci0 + c >= Cin) break;
            float in[2][2][3];  // [dz][dy][column xx + 0..2]
#pragma unroll
            for (int i = 0; i < 12; ++i) (&in[0][0][0])[i] = (&nxt[0][0][0])[i];
            if (c + 1 < CK) load_inputs(ci0 + c + 1, nxt);
            const float *wc = s_w + c * 27 * CT;
#pragma unroll
            for (int p = 0; p < 8; ++p) {           // output parity (pz,py,px)
                const int pz = p >> 2, py = (p >> 1) & 1, px = p & 1;
#pragma unroll
                for (int n = 0; n < 8; ++n) {       // input offset (dz,dy,dx)
                    const int dz = n >> 2, dy = (n >> 1) & 1, dx = n & 1;
                    if ((dz && !pz) || (dy && !py) || (dx && !px)) continue;  // even outputs only see offset 0
                    const int kd = pz ? (dz ? 0 : 2) : 1;
                    const int kh = py ? (dy ? 0 : 2) : 1;
                    const int kw = px ? (dx ? 0 : 2) : 1;
                    const float4 wt = *reinterpret_cast<const float4 *>(wc + ((kd * 3 + kh) * 3 + kw) * CT);
                    const float2 w01 = make_float2(wt.x, wt.y), w23 = make_float2(wt.z, wt.w);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float a = in[dz][dy][j + dx];
                        const float2 a2 = make_float2(a, a);
                        acc[j][p][0] = __ffma2_rn(a2, w01, acc[j][p][0]);
                        acc[j][p][1] = __ffma2_rn(a2, w23, acc[j][p][1]);
                    }
                }
            }
        }
    }
    if (!live) return;
    if (skip_staged) asm volatile("cp.async.wait_group 0;" ::: "memory");  // own copies only: no barrier needed
#pragma unroll
    for (int q = 0; q < CT; ++q) {
        if (co0 + q >= Cout) break;
        const float sh = __ldg(shift + co0 + q);
#pragma unroll
        for (int pzy = 0; pzy < 4; ++pzy) {
            const int pz = pzy >> 1, py = pzy & 1;
            const size_t o = ((size_t)b * Cout + co0 + q) * out_cs + ((size_t)(2 * z + pz) * Ho + 2 * yy + py) * Wo + 2 * xx;
            float v[4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int px = 0; px < 2; ++px) {
                    const float2 a = acc[j][pzy * 2 + px][q >> 1];
                    float t = ((q & 1) ? a.y : a.x) + sh;
                    if (relu) t = fmaxf(t, 0.f);
                    v[2 * j + px] = t;
                }
            if (vec4) {
                if (skip) {  // skip + relu(bn(convT))   (mvsnet.py:69-71)
                    const float4 sk = s_skip[q * 4 + pzy][tid];
                    v[0] += sk.x; v[1] += sk.y; v[2] += sk.z; v[3] += sk.w;
                }
                *reinterpret_cast<float4 *>(y + o) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j == 1 && !hx1) break;
                    float v0 = v[2 * j], v1 = v[2 * j + 1];
                    if (skip) {
                        const float2 sk = __ldg(reinterpret_cast<const float2 *>(skip + o + 2 * j));
                        v0 += sk.x;
                        v1 += sk.y;
                    }
                    *reinterpret_cast<float2 *>(y + o + 2 * j) = make_float2(v0, v1);
                }
            }
        }
    }
}

int conv3d_fp32(const float *x, const float *w, const float *shift, int relu, float *y, int B, int Cin, int Cout, int D,
                int H, int W, int stride, cudaStream_t st) {
    if (stride == 1) {
        if (Cout < 8) return launch_conv<1, 4, 4, 8, 1>(x, w, shift, relu, y, B, Cin, Cout, D, H, W, st);
        return launch_conv<1, 4, 4, 8, 8>(x, w, shift, relu, y, B, Cin, Cout, D, H, W, st);
    }
    if (Cout < 8) return launch_conv<2, 2, 2, 8, 1>(x, w, shift, relu, y, B, Cin, Cout, D, H, W, st);
    return launch_conv<2, 2, 2, 8, 8>(x, w, shift, relu, y, B, Cin, Cout, D, H, W, st);
}

int convT3d_fp32(const float *x, const float *w, const float *shift, int relu, const float *skip, float *y, int B,
                 int Cin, int Cout, int D, int H, int W, cudaStream_t st) {
    MVS_REQUIRE(D <= 65535, "conv_transpose3d: D too large");
    if (Cout < 8) {
        const int cg = Cout;
        MVS_REQUIRE((long long)B * cg <= 65535, "conv_transpose3d: grid too large");
        convT3d_fp32_kernel<1, 16><<<dim3(cdiv((long long)H * W, 128), D, B * cg), 128, 0, st>>>(x, w, shift, relu, skip, y,
                                                                                               Cin, Cout, D, H, W);
    } else {
        const int cg = cdiv(Cout, 4);
        MVS_REQUIRE((long long)B * cg <= 65535, "conv_transpose3d: grid too large");
        convT3d_fp32_pair_kernel<16><<<dim3(cdiv((long long)H * ((W + 1) / 2), 128), D, B * cg), 128, 0, st>>>(x, w, shift, relu, skip,
                                                                                                          y, Cin, Cout, D, H, W);
    }
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_conv3d_bn_relu(const float *x, const float *w, const float *shift, int relu, float *y, int B, int Cin,
                                  int Cout, int D, int H, int W, int stride, void *stream) {
    MVS_REQUIRE(x && w && shift && y, "null pointer argument");
    MVS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "bad shape");
    MVS_REQUIRE(stride == 1 || stride == 2, "conv3d: stride must be 1 or 2, got %d", stride);
    return conv3d_fp32(x, w, shift, relu, y, B, Cin, Cout, D, H, W, stride, (cudaStream_t)stream);
}

extern "C" int mvs_conv_transpose3d_bn_relu(const float *x, const float *w, const float *shift, int relu,
                                            const float *skip, float *y, int B, int Cin, int Cout, int D, int H, int W,
                                            void *stream) {
    MVS_REQUIRE(x && w && shift && y, "null pointer argument");
    MVS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "bad shape");
    return convT3d_fp32(x, w, shift, relu, skip, y, B, Cin, Cout, D, H, W, (cudaStream_t)stream);
}
