// Depth-axis softmax + depth expectation + 4-plane photometric confidence in one kernel.
//
// Replaces, for the reference:
//   models/mvsnet.py:192-193   F.softmax(cost_reg.squeeze(1), dim=1)
//   models/module.py:144-147   depth_regression  (sum_d p * depth_values)
//   models/mvsnet.py:214-218   pad + avg_pool3d(4,1,1)*4 + depth_regression(arange).long() + gather
// which in the reference are ~8 ATen kernels and five [B,D,H,W] temporaries.
//
// HBM-bound: the logits are read exactly once (4*B*D*H*W bytes) and 8*B*H*W bytes are written.
// A CTA owns 32 consecutive pixels; its 8 warps split the depth axis into 8 slices so every global
// load is a full 128-byte line (lanes run along x, the contiguous axis) and 8 independent load
// streams per pixel hide HBM latency.  The logits tile is kept in shared memory between the max
// pass and the exp/sum pass, exactly the reference's  exp(l - max) / sum  formulation.
//
// Two kernels with the same arithmetic (bit-identical results): the TMA-pipelined one below for the shapes TMA can
// describe (H*W % 4 == 0, D <= 256), persistent CTAs that prefetch the next pixel group's [D][64] logits tile with one
// bulk tensor copy while the current one is reduced; and the direct one (any shape).
#include <float.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace mvs {

constexpr int kSlices = 8;
constexpr int kTailPx = 64;  // pixels per group of the pipelined kernel

// ------------------------------------------------------------------------------------------------
// Pipelined variant.  grid = 2 CTAs per SM (persistent), block = 64 pixels x 8 depth slices.  Per group: one TMA box
// {64 pixels, D planes} (zero-filled beyond H*W) lands in one of two stages; max pass, exp/sum pass (e written back to
// the tile for the confidence window and the optional probability output), finalisation by slice 0.  The sums use the
// same slice boundaries and the same fixed reduction order as the direct kernel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTailPx * kSlices, 2)
softmax_depth_conf_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ depth_values,
                              float *__restrict__ depth, float *__restrict__ conf, float *__restrict__ prob, int D, int HW,
                              int groups_per_b, int n_groups) {
    extern __shared__ __align__(128) float s_stage[];  // [2][D][64]
    __shared__ float s_red[4][kSlices][kTailPx];
    __shared__ __align__(8) uint64_t s_bar[2];
    const int px = threadIdx.x & (kTailPx - 1), slice = threadIdx.x / kTailPx;
    const int dq = (D + kSlices - 1) / kSlices;
    const int d0 = slice * dq, d1 = min(D, d0 + dq);
    const uint32_t stage_bytes = (uint32_t)D * kTailPx * 4u;
    const uint32_t bar0 = ptx::smem_u32(&s_bar[0]), st0 = ptx::smem_u32(s_stage);
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar0, 1);
        ptx::mbar_init(bar0 + 8, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&tmap);
    }
    __syncthreads();
    // launched with programmatic stream serialization: CTAs may start while the producer of `logits` drains
    ptx::pdl_wait();
    int g = blockIdx.x;
    if (threadIdx.x == 0 && g < n_groups) {
        ptx::mbar_arrive_expect_tx(bar0, stage_bytes);
        ptx::tma_load_3d(st0, &tmap, bar0, (g % groups_per_b) * kTailPx, 0, g / groups_per_b);
    }
    uint32_t par = 0;  // bit s: parity of stage s
    for (int k = 0; g < n_groups; g += gridDim.x, ++k) {
        const int s = k & 1;
        const int gn = g + gridDim.x;
        if (threadIdx.x == 0 && gn < n_groups) {  // prefetch the next group (its stage was released by the barrier below)
            ptx::mbar_arrive_expect_tx(bar0 + 8 * (s ^ 1), stage_bytes);
            ptx::tma_load_3d(st0 + (uint32_t)(s ^ 1) * stage_bytes, &tmap, bar0 + 8 * (s ^ 1), (gn % groups_per_b) * kTailPx, 0,
                             gn / groups_per_b);
        }
        ptx::mbar_wait(bar0 + 8 * s, (par >> s) & 1u);
        par ^= 1u << s;
        float *tile = s_stage + (size_t)s * D * kTailPx;
        const int b = g / groups_per_b;
        const int pix = (g % groups_per_b) * kTailPx + px;
        const bool live = pix < HW;
        const float *dv = depth_values + (size_t)b * D;

        float m = -FLT_MAX;
#pragma unroll 8
        for (int d = d0; d < d1; ++d) m = fmaxf(m, tile[d * kTailPx + px]);
        s_red[0][slice][px] = m;
        __syncthreads();
        float M = s_red[0][0][px];
#pragma unroll
        for (int q = 1; q < kSlices; ++q) M = fmaxf(M, s_red[0][q][px]);

        float se = 0.f, sd = 0.f, si = 0.f;
#pragma unroll 4
        for (int d = d0; d < d1; ++d) {
            const float e = expf(tile[d * kTailPx + px] - M);
            tile[d * kTailPx + px] = e;
            se += e;
            sd = fmaf(e, __ldg(dv + d), sd);
            si = fmaf(e, (float)d, si);
        }
        s_red[1][slice][px] = se;
        s_red[2][slice][px] = sd;
        s_red[3][slice][px] = si;
        __syncthreads();
        float sum = 0.f, sumd = 0.f, sumi = 0.f;
#pragma unroll
        for (int q = 0; q < kSlices; ++q) {  // fixed order: deterministic
            sum += s_red[1][q][px];
            sumd += s_red[2][q][px];
            sumi += s_red[3][q][px];
        }
        if (slice == 0 && live) {
            const float idxf = sumi / sum;  // sum_d p[d] * d
            int i = (int)idxf;              // .long() truncation (mvsnet.py:217)
            i = min(max(i, 0), D - 1);
            float c4 = 0.f;
#pragma unroll
            for (int kq = -1; kq <= 2; ++kq) {  // p[i-1] + p[i] + p[i+1] + p[i+2], zero padded (mvsnet.py:216)
                const int kk = i + kq;
                if (kk >= 0 && kk < D) c4 += tile[kk * kTailPx + px] / sum;
            }
            depth[(size_t)b * HW + pix] = sumd / sum;
            conf[(size_t)b * HW + pix] = c4;
        }
        if (prob != nullptr && live) {
            float *pp = prob + (size_t)b * D * HW + pix;
            for (int d = d0; d < d1; ++d) pp[(size_t)d * HW] = tile[d * kTailPx + px] / sum;
        }
        ptx::fence_proxy_async_smem();  // this thread's writes of e into the tile precede the TMA refill of the stage
        __syncthreads();  // every read of this stage (and of s_red) is done: the next iteration may refill it
    }
}

template <bool CACHE>
__global__ void __launch_bounds__(32 * kSlices)
softmax_depth_conf_kernel(const float *__restrict__ logits, const float *__restrict__ depth_values,
                          float *__restrict__ depth, float *__restrict__ conf, float *__restrict__ prob, int D,
                          int HW) {
    extern __shared__ float s_tile[];  // CACHE: [D][32] logits
    // launched with programmatic stream serialization: CTAs may start while the producer of `logits` drains
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __shared__ float s_red[4][kSlices][32];

    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int pix = blockIdx.x * 32 + lane;
    const bool live = pix < HW;
    const int dq = (D + kSlices - 1) / kSlices;
    const int d0 = slice * dq, d1 = min(D, d0 + dq);
    const float *lp = logits + (size_t)b * D * HW + (live ? pix : 0);
    const float *dv = depth_values + (size_t)b * D;

    // pass 1: max over depth
    float m = -FLT_MAX;
#pragma unroll 24  // D = 192 with 8 slices: a thread's 24 loads all in flight at once
    for (int d = d0; d < d1; ++d) {
        const float l = live ? __ldcs(lp + (size_t)d * HW) : 0.f;
        if (CACHE) s_tile[d * 32 + lane] = l;
        m = fmaxf(m, l);
    }
    s_red[0][slice][lane] = m;
    __syncthreads();
    float M = s_red[0][0][lane];
#pragma unroll
    for (int s = 1; s < kSlices; ++s) M = fmaxf(M, s_red[0][s][lane]);

    // pass 2: e = exp(l - M); partial sums of e, e*depth, e*index
    float se = 0.f, sd = 0.f, si = 0.f;
#pragma unroll 4
    for (int d = d0; d < d1; ++d) {
        const float l = CACHE ? s_tile[d * 32 + lane] : (live ? __ldg(lp + (size_t)d * HW) : 0.f);
        const float e = expf(l - M);
        if (CACHE) s_tile[d * 32 + lane] = e;
        se += e;
        sd = fmaf(e, __ldg(dv + d), sd);
        si = fmaf(e, (float)d, si);
    }
    s_red[1][slice][lane] = se;
    s_red[2][slice][lane] = sd;
    s_red[3][slice][lane] = si;
    __syncthreads();
    float sum = 0.f, sumd = 0.f, sumi = 0.f;
#pragma unroll
    for (int s = 0; s < kSlices; ++s) {  // fixed order: deterministic
        sum += s_red[1][s][lane];
        sumd += s_red[2][s][lane];
        sumi += s_red[3][s][lane];
    }

    if (slice == 0 && live) {
        const float idxf = sumi / sum;  // sum_d p[d] * d
        int i = (int)idxf;              // .long() truncation (mvsnet.py:217)
        i = min(max(i, 0), D - 1);
        float c4 = 0.f;
#pragma unroll
        for (int k = -1; k <= 2; ++k) {  // p[i-1] + p[i] + p[i+1] + p[i+2], zero padded (mvsnet.py:216)
            const int kk = i + k;
            if (kk >= 0 && kk < D) {
                const float e = CACHE ? s_tile[kk * 32 + lane] : expf(__ldg(lp + (size_t)kk * HW) - M);
                c4 += e / sum;
            }
        }
        depth[(size_t)b * HW + pix] = sumd / sum;
        conf[(size_t)b * HW + pix] = c4;
    }
    if (prob != nullptr && live) {
        float *pp = prob + (size_t)b * D * HW + pix;
        for (int d = d0; d < d1; ++d) {
            const float e = CACHE ? s_tile[d * 32 + lane] : expf(__ldg(lp + (size_t)d * HW) - M);
            pp[(size_t)d * HW] = e / sum;
        }
    }
}

__global__ void depth_regression_kernel(const float *__restrict__ p, const float *__restrict__ depth_values,
                                        int dv_stride, float *__restrict__ out, int D, int HW) {
    const int b = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const float *pp = p + (size_t)b * D * HW + pix;
    const float *dv = depth_values + (size_t)b * dv_stride;
    float acc = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) acc = fmaf(__ldg(pp + (size_t)d * HW), __ldg(dv + d), acc);
    out[(size_t)b * HW + pix] = acc;
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_softmax_depth_conf(const float *logits, const float *depth_values, float *depth, float *conf,
                                      float *prob, int B, int D, int H, int W, void *stream) {
    MVS_REQUIRE(logits && depth_values && depth && conf, "null pointer argument");
    MVS_REQUIRE(B > 0 && B <= 65535 && D > 0 && H > 0 && W > 0, "bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    // ---- TMA-pipelined kernel when a tensor map can describe the logits
    const size_t smem_tma = 2 * (size_t)D * kTailPx * sizeof(float);
    if (HW % 4 == 0 && D <= 256 && D >= kSlices && ((uintptr_t)logits & 15) == 0 && smem_tma <= 100 * 1024) {
        tmap_encode_fn enc = get_tmap_encode();
        MVS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
        CUtensorMap tmap;
        cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)D, (cuuint64_t)B};
        cuuint64_t gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)HW * D * 4};
        cuuint32_t box[3] = {(cuuint32_t)kTailPx, (cuuint32_t)D, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(logits), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return set_error(MVS_ERR_CUDA, "cuTensorMapEncodeTiled (logits) failed (%d)", (int)cr);
        MVS_CUDA(cudaFuncSetAttribute(softmax_depth_conf_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        int dev = 0, sms = 148;
        MVS_CUDA(cudaGetDevice(&dev));
        MVS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int groups_per_b = cdiv(HW, kTailPx);
        const long long n_groups = (long long)groups_per_b * B;
        MVS_REQUIRE(n_groups < (1LL << 31), "too many pixel groups");
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)std::min<long long>(n_groups, 2LL * sms));
        cfg.blockDim = dim3(kTailPx * kSlices);
        cfg.dynamicSmemBytes = smem_tma;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MVS_CUDA(cudaLaunchKernelEx(&cfg, softmax_depth_conf_tma_kernel, tmap, depth_values, depth, conf, prob, D, HW, groups_per_b,
                                    (int)n_groups));
        MVS_LAUNCH_CHECK(1);
        return MVS_OK;
    }
    dim3 grid(cdiv(HW, 32), B);
    const size_t smem = (size_t)D * 32 * sizeof(float);
    if (smem <= 96 * 1024) {
        if (smem > 48 * 1024)
            MVS_CUDA(cudaFuncSetAttribute(softmax_depth_conf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          96 * 1024));  // one value for every caller (the attribute is per function, not per launch)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(32 * kSlices);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MVS_CUDA(cudaLaunchKernelEx(&cfg, softmax_depth_conf_kernel<true>, logits, depth_values, depth, conf, prob, D, HW));
    } else {
        softmax_depth_conf_kernel<false><<<grid, 32 * kSlices, 0, st>>>(logits, depth_values, depth, conf, prob, D, HW);
    }
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

extern "C" int mvs_depth_regression(const float *p, const float *depth_values, int dv_batch_stride, float *out, int B,
                                    int D, int H, int W, void *stream) {
    MVS_REQUIRE(p && depth_values && out, "null pointer argument");
    MVS_REQUIRE(B > 0 && B <= 65535 && D > 0 && H > 0 && W > 0, "bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
    MVS_REQUIRE(dv_batch_stride == 0 || dv_batch_stride >= D, "depth_values batch stride must be 0 or >= D");
    const int HW = H * W;
    depth_regression_kernel<<<dim3(cdiv(HW, 128), B), 128, 0, (cudaStream_t)stream>>>(p, depth_values, dv_batch_stride,
                                                                                      out, D, HW);
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}
