// FeatureNet building block and FeatureNet itself, strict-precision (fp32 FMA) path on the CUDA cores.
//
// Replaces, for the reference:
//   models/module.py:8-15    ConvBnReLU = nn.Conv2d(k, stride, pad, bias=False) + BatchNorm2d + ReLU
//   models/mvsnet.py:10-30   FeatureNet: conv0 3->8, conv1 8->8, conv2 8->16 (k5 s2), conv3 / conv4 16->16,
//                            conv5 16->32 (k5 s2), conv6 32->32, feature = nn.Conv2d(32, 32, 3, 1, 1) (bias, no BN / ReLU)
// Eval-mode BatchNorm is folded into weights / shift by the caller (models/mvsnet.py FeatureNet.folded_native).
// The tensor-core form of the same network is featurenet_tc in conv3d_tc.cu; before this file the strict mode ran
// FeatureNet on cuDNN without TF32 (8.9 ms of its 33 ms per depth map at the DTU shape).
//
// Same structure as conv3d_fp32_tma_kernel: a CTA owns 32 x 32 outputs x 8 output channels, a thread 4 consecutive x
// outputs x 8 channels in registers; the input halo tile of a chunk of CK channels arrives by TMA (4-D box x, y, channel,
// image; zero fill outside the image = the convolution's padding), two stages; the chunk's weights by cp.async.
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mvs {

namespace {

constexpr int kTY = 32;        // output rows per CTA
constexpr int kCoutT = 8;      // output channels per CTA
constexpr int kThreads2d = 8 * kTY;

template <int K, int S>
struct Tile2d {
    static_assert((K == 3 && S == 1) || (K == 5 && S == 2), "FeatureNet layers: k3 s1 p1 or k5 s2 p2");
    static constexpr int PAD = K / 2;
    // The innermost start coordinate of a TMA box must be a multiple of 16 bytes.  k3 s1: the output tile starts at
    // 32 i - 3, its inputs at 32 i - 4, a thread's six inputs are an aligned float4 + float2 of its row (and its stores are
    // scalar).  k5 s2: outputs at 32 i, inputs needed from 64 i - 2, box from 64 i - 4: a thread's eleven inputs start at
    // column 2 of its 8-column span (float2, float4, float4, scalar).
    static constexpr int OUT_SHIFT = (S == 1) ? 3 : 0;
    static constexpr int X_LEAD = 4;                               // box start = S * ox0 - X_LEAD
    static constexpr int XO = X_LEAD - PAD - OUT_SHIFT * S;        // first input of a thread, relative to tx * 4 * S
    static constexpr int IY = (kTY - 1) * S + K;
    static constexpr int IX = XO + 31 * S + K;
    static constexpr int IXP = (IX + 3) / 4 * 4;
    static constexpr int PER_CH = IY * IXP;
    static constexpr int NIN = 3 * S + K;                          // inputs of a thread per row
};
static_assert(Tile2d<3, 1>::XO == 0 && Tile2d<3, 1>::IXP == 36, "k3 s1 tile");
static_assert(Tile2d<5, 2>::XO == 2 && Tile2d<5, 2>::IXP == 72, "k5 s2 tile");

__device__ __forceinline__ void cp_async4_zfill(float *dst_smem, const float *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src),
                 "r"(src_bytes)
                 : "memory");
}

template <int K, int S, int CK>
__global__ void __launch_bounds__(kThreads2d)
conv2d_fp32_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ w, const float *__restrict__ shift,
                       int relu, float *__restrict__ y, int Cin, int Cout, int Ho, int Wo, int tiles_x) {
    using T = Tile2d<K, S>;
    constexpr int kBox = CK * T::PER_CH;          // floats a TMA box delivers
    constexpr int kTile = (kBox + 31) / 32 * 32;  // floats per stage: 128-byte aligned TMA destinations
    constexpr int kW = CK * K * K * kCoutT;
    extern __shared__ __align__(128) float smem2d[];
    float *s_in = smem2d;                                           // [2][CK][IY][IXP]
    float *s_w = s_in + 2 * kTile;                                  // [2][CK][K][K][8]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_w + 2 * kW);   // [2]

    const int tid = threadIdx.x;
    const int tx = tid & 7, ty = tid >> 3;
    const int cgroups = (Cout + kCoutT - 1) / kCoutT;
    const int n = blockIdx.z / cgroups;
    const int co0 = (blockIdx.z % cgroups) * kCoutT;
    const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * kTY;
    const int nchunks = (Cin + CK - 1) / CK;
    const uint32_t bar0 = ptx::smem_u32(s_bar), in0 = ptx::smem_u32(s_in);
    const CUtensorMap *const tm = &tmap;

    if (tid == 0) {
        // a stage is complete when the TMA box has landed (one arrival + its bytes) and every thread's weight copies have
        // (cp.async.mbarrier.arrive.noinc: one arrival per thread when its earlier cp.async are done)
        ptx::mbar_init(bar0, 1 + kThreads2d);
        ptx::mbar_init(bar0 + 8, 1 + kThreads2d);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(tm);
    }
    __syncthreads();

    auto issue = [=](int k) {
        const int st = k & 1, ci0 = k * CK;
        if (tid == 0) {
            ptx::fence_proxy_async_smem();  // the stage was read through the generic proxy two chunks ago
            ptx::mbar_arrive_expect_tx(bar0 + 8 * st, (uint32_t)(kBox * 4));
            ptx::tma_load_4d(in0 + (uint32_t)(st * kTile * 4), tm, bar0 + 8 * st, ox0 * S - T::X_LEAD, oy0 * S - T::PAD, ci0, n);
        }
        float *dw = s_w + st * kW;
        for (int idx = tid; idx < kW; idx += kThreads2d) {
            const int co = idx % kCoutT;
            const int tap = (idx / kCoutT) % (K * K);
            const int c = idx / (kCoutT * K * K);
            const bool ok = co0 + co < Cout && ci0 + c < Cin;
            cp_async4_zfill(dw + idx, w + (ok ? ((size_t)(co0 + co) * Cin + ci0 + c) * (K * K) + tap : 0), ok ? 4 : 0);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar0 + 8 * st) : "memory");
    };

    float2 acc[4][kCoutT / 2];  // packed pairs of output channels (FFMA2: two IEEE fmas per issue slot)
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int q = 0; q < kCoutT / 2; ++q) acc[o][q] = make_float2(0.f, 0.f);

    issue(0);
    for (int k = 0; k < nchunks; ++k) {
        const int st = k & 1;
        if (k + 1 < nchunks) issue(k + 1);  // into the other stage: its readers finished behind the barrier that ended chunk k - 1
        ptx::mbar_wait(bar0 + 8 * st, (uint32_t)((k >> 1) & 1));  // tile and weights of this chunk are in shared memory
        const float *tin = s_in + st * kTile, *tw = s_w + st * kW;
#pragma unroll 1
        for (int c = 0; c < CK; ++c) {
#pragma unroll
            for (int kh = 0; kh < K; ++kh) {
                const float *row = tin + (c * T::IY + ty * S + kh) * T::IXP + tx * 4 * S;
                float in[T::NIN];
                if (S == 1) {
                    const float4 a = *reinterpret_cast<const float4 *>(row);
                    const float2 e = *reinterpret_cast<const float2 *>(row + 4);
                    in[0] = a.x; in[1] = a.y; in[2] = a.z; in[3] = a.w; in[4] = e.x; in[5] = e.y;
                } else {
                    const float2 p = *reinterpret_cast<const float2 *>(row + 2);
                    const float4 a = *reinterpret_cast<const float4 *>(row + 4);
                    const float4 e = *reinterpret_cast<const float4 *>(row + 8);
                    in[0] = p.x; in[1] = p.y;
                    in[2] = a.x; in[3] = a.y; in[4] = a.z; in[5] = a.w;
                    in[6] = e.x; in[7] = e.y; in[8] = e.z; in[9] = e.w;
                    in[10] = row[12];
                }
                const float *wp = tw + (c * K + kh) * K * kCoutT;
#pragma unroll
                for (int kw = 0; kw < K; ++kw) {
                    const float4 w0 = *reinterpret_cast<const float4 *>(wp + kw * kCoutT);
                    const float4 w1 = *reinterpret_cast<const float4 *>(wp + kw * kCoutT + 4);
                    const float2 wr[kCoutT / 2] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                                   make_float2(w1.z, w1.w)};
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const float2 a2 = make_float2(in[o * S + kw], in[o * S + kw]);
#pragma unroll
                        for (int q = 0; q < kCoutT / 2; ++q) acc[o][q] = __ffma2_rn(a2, wr[q], acc[o][q]);
                    }
                }
            }
        }
        __syncthreads();  // the stage may be refilled
    }

    const int oy = oy0 + ty, ox = ox0 - T::OUT_SHIFT + tx * 4;
    if (oy >= Ho || ox >= Wo) return;
#pragma unroll
    for (int q = 0; q < kCoutT; ++q) {
        if (co0 + q >= Cout) break;
        const float sh = __ldg(shift + co0 + q);
        float v[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            v[o] = ((q & 1) ? acc[o][q >> 1].y : acc[o][q >> 1].x) + sh;
            if (relu) v[o] = fmaxf(v[o], 0.f);
        }
        float *op = y + (((size_t)n * Cout + co0 + q) * Ho + oy) * Wo + ox;
        if (S == 2 && (Wo & 3) == 0) {
            *reinterpret_cast<float4 *>(op) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (ox + o >= 0 && ox + o < Wo) op[o] = v[o];
        }
    }
}

template <int K, int S, int CK>
int launch_conv2d(const float *x, const float *w, const float *shift, int relu, float *y, int N, int Cin, int Cout, int H, int W,
                  cudaStream_t st) {
    using T = Tile2d<K, S>;
    const int Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1;
    const int tiles_x = cdiv(Wo + T::OUT_SHIFT, 32), tiles_y = cdiv(Ho, kTY);
    const int cgroups = cdiv(Cout, kCoutT);
    MVS_REQUIRE(tiles_y <= 65535 && (long long)N * cgroups <= 65535, "conv2d: grid too large");
    tmap_encode_fn enc = get_tmap_encode();
    MVS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap tmap;
    const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Cin, (cuuint64_t)N};
    const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)Cin * H * W * 4};
    const cuuint32_t box[4] = {(cuuint32_t)T::IXP, (cuuint32_t)T::IY, (cuuint32_t)CK, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return set_error(MVS_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 conv2d input) failed (%d)", (int)cr);
    const size_t smem = (size_t)2 * ((CK * T::PER_CH + 31) / 32 * 32 + CK * K * K * kCoutT) * sizeof(float) + 16;
    auto kern = conv2d_fp32_tma_kernel<K, S, CK>;
    MVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(tiles_x, tiles_y, N * cgroups), kThreads2d, smem, st>>>(tmap, w, shift, relu, y, Cin, Cout, Ho, Wo, tiles_x);
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

}  // namespace

// One ConvBnReLU (BN folded): x [N,Cin,H,W] -> y [N,Cout,H',W'], ksize 3 / stride 1 / pad 1 or ksize 5 / stride 2 / pad 2.
// W % 4 == 0 and a 16-byte aligned x (tensor-map strides).  Asynchronous on st.
int conv2d_fp32(const float *x, const float *w, const float *shift, int relu, float *y, int N, int Cin, int Cout, int H, int W,
                int ksize, int stride, cudaStream_t st) {
    MVS_REQUIRE((ksize == 3 && stride == 1) || (ksize == 5 && stride == 2), "conv2d: ksize/stride %d/%d (3/1 or 5/2)", ksize, stride);
    MVS_REQUIRE((W & 3) == 0 && ((uintptr_t)x & 15) == 0, "conv2d: W = %d must be a multiple of 4 and x 16-byte aligned", W);
    if (ksize == 3) {
        if (Cin <= 3) return launch_conv2d<3, 1, 3>(x, w, shift, relu, y, N, Cin, Cout, H, W, st);
        return launch_conv2d<3, 1, 4>(x, w, shift, relu, y, N, Cin, Cout, H, W, st);
    }
    return launch_conv2d<5, 2, 2>(x, w, shift, relu, y, N, Cin, Cout, H, W, st);
}

static const int kFnCin[MVS_FEATURENET_LAYERS] = {3, 8, 8, 16, 16, 16, 32, 32};
static const int kFnCout[MVS_FEATURENET_LAYERS] = {8, 8, 16, 16, 16, 32, 32, 32};
static const int kFnK[MVS_FEATURENET_LAYERS] = {3, 3, 5, 3, 3, 5, 3, 3};

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_conv2d_bn_relu(const float *x, const float *w, const float *shift, int relu, float *y, int N, int Cin,
                                  int Cout, int H, int W, int ksize, int stride, void *stream) {
    MVS_REQUIRE(x && w && shift && y, "null pointer argument");
    MVS_REQUIRE(N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "bad shape");
    return conv2d_fp32(x, w, shift, relu, y, N, Cin, Cout, H, W, ksize, stride, (cudaStream_t)stream);
}

// two ping-pong activation buffers of the largest layer output (8 channels at full resolution)
extern "C" size_t mvs_featurenet_workspace_bytes(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0 || (H & 3) || (W & 15)) return 0;
    return (size_t)2 * N * 8 * H * W * sizeof(float);
}

extern "C" int mvs_featurenet_fwd(const float *imgs, const mvs_featurenet_params *params, float *fea, void *workspace, int N,
                                  int H, int W, void *stream) {
    MVS_REQUIRE(imgs && params && fea && workspace, "null pointer argument");
    MVS_REQUIRE(N > 0 && H > 0 && W > 0 && (H & 3) == 0 && (W & 15) == 0, "featurenet: H %% 4 == 0 and W %% 16 == 0 required, got %dx%d", H, W);
    cudaStream_t st = (cudaStream_t)stream;
    float *buf[2] = {(float *)workspace, (float *)workspace + (size_t)N * 8 * H * W};
    const float *in = imgs;
    int h = H, w = W;
    for (int l = 0; l < MVS_FEATURENET_LAYERS; ++l) {
        MVS_REQUIRE(params->w[l] && params->shift[l], "featurenet: layer %d parameters missing", l);
        const int stride = kFnK[l] == 5 ? 2 : 1;
        float *out = (l == MVS_FEATURENET_LAYERS - 1) ? fea : buf[l & 1];
        if (int rc = conv2d_fp32(in, params->w[l], params->shift[l], l != MVS_FEATURENET_LAYERS - 1, out, N, kFnCin[l], kFnCout[l], h,
                                 w, kFnK[l], stride, st))
            return rc;
        in = out;
        h = (h - 1) / stride + 1;
        w = (w - 1) / stride + 1;
    }
    return MVS_OK;
}
