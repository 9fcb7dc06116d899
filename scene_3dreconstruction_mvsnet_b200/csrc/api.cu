// C-ABI glue of libmvsnet_b200.so: error reporting, launch accounting, the CostRegNet driver and the
// host-buffer entry point.  See include/mvsnet_b200.h for the contract.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mvs {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int conv3d_fp32(const float *x, const float *w, const float *shift, int relu, float *y, int B, int Cin, int Cout, int D,
                int H, int W, int stride, cudaStream_t st);
int convT3d_fp32(const float *x, const float *w, const float *shift, int relu, const float *skip, float *y, int B,
                 int Cin, int Cout, int D, int H, int W, cudaStream_t st);
int costreg_tc(const float *volume, const void *volume_cp8, const mvs_costreg_params *p, float *logits, void *workspace,
               int B, int D, int H, int W, cudaStream_t st);
int warp_variance_cp8(const float *fea, const float *proj, const float *depth_values, void *vol_cp8, void *workspace,
                      int B, int V, int D, int H, int W, cudaStream_t st);
size_t costreg_tc_workspace_bytes(int B, int D, int H, int W);
int warp_variance_cp8_rcp8(const void *tex16, const float *proj, const float *depth_values, void *vol_cp8, void *workspace,
                           int B, int V, int D, int H, int W, cudaStream_t st);
int warp_variance_cp8_f16(const void *fea16, const float *proj, const float *depth_values, void *vol_cp8, void *workspace,
                          int B, int V, int D, int H, int W, cudaStream_t st);
int warp_variance_cp8_pool(const void *pool16, int n_pool, const int *view_ids_host, const float *proj,
                           const float *depth_values, void *vol_cp8, void *workspace, int B, int V, int D, int H, int W,
                           cudaStream_t st);

// layer table of CostRegNet (mvsnet.py:36-62): {Cin, Cout}; order conv0..conv6, conv7, conv9, conv11, prob
static const int kLayerCin[MVS_COSTREG_LAYERS] = {32, 8, 16, 16, 32, 32, 64, 64, 32, 16, 8};
static const int kLayerCout[MVS_COSTREG_LAYERS] = {8, 16, 16, 32, 32, 64, 64, 32, 16, 8, 1};

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_abi_version(void) { return MVSNET_B200_ABI_VERSION; }
extern "C" const char *mvs_last_error(void) { return g_err; }
extern "C" uint64_t mvs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char *mvs_arch(void) { return "sm_100a"; }

// ---------------------------------------------------------------------------------------------
// CostRegNet.forward (mvsnet.py:64-73), fp32 path: 11 fused conv(+BN+ReLU+skip) launches.
// workspace (floats, per batch element, N0 = D*H*W):
//   c0 8N0 | c1 2N0 | c2 2N0 | c3 N0/2 | c4 N0/2 | c5 N0/8 | c6 N0/8 | u7 N0/2 | u9 2N0 | u11 8N0
// ---------------------------------------------------------------------------------------------
static size_t costreg_fp32_floats(int B, int D, int H, int W) {
    const size_t n0 = (size_t)D * H * W;
    return (size_t)B * (8 * n0 + 2 * n0 + 2 * n0 + n0 / 2 + n0 / 2 + n0 / 8 + n0 / 8 + n0 / 2 + 2 * n0 + 8 * n0);
}

extern "C" size_t mvs_costreg_workspace_bytes(int B, int D, int H, int W, int precision) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || (D % 8) || (H % 8) || (W % 8)) return 0;
    if (precision == MVS_PRECISION_BF16) return costreg_tc_workspace_bytes(B, D, H, W);
    return costreg_fp32_floats(B, D, H, W) * sizeof(float);
}

extern "C" int mvs_costreg_fwd(const float *volume, const mvs_costreg_params *p, float *logits, void *workspace, int B,
                               int D, int H, int W, int precision, void *stream) {
    MVS_REQUIRE(volume && p && logits && workspace, "null pointer argument");
    MVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
    // three stride-2 stages then output_padding=1 + skip add: the reference fails with a size mismatch otherwise
    MVS_REQUIRE(D % 8 == 0 && H % 8 == 0 && W % 8 == 0,
                "CostRegNet needs D, H, W divisible by 8 (got D=%d H=%d W=%d): conv4 + conv7(x) would not match", D, H,
                W);
    for (int i = 0; i < MVS_COSTREG_LAYERS; ++i)
        MVS_REQUIRE(p->w[i] && p->shift[i], "costreg params: layer %d has a null pointer", i);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == MVS_PRECISION_BF16) return costreg_tc(volume, nullptr, p, logits, workspace, B, D, H, W, st);
    MVS_REQUIRE(precision == MVS_PRECISION_FP32, "unknown precision %d", precision);

    const size_t n0 = (size_t)D * H * W * B;
    float *ws = (float *)workspace;
    float *c0 = ws;            ws += 8 * n0;
    float *c1 = ws;            ws += 2 * n0;
    float *c2 = ws;            ws += 2 * n0;
    float *c3 = ws;            ws += n0 / 2;
    float *c4 = ws;            ws += n0 / 2;
    float *c5 = ws;            ws += n0 / 8;
    float *c6 = ws;            ws += n0 / 8;
    float *u7 = ws;            ws += n0 / 2;
    float *u9 = ws;            ws += 2 * n0;
    float *u11 = ws;
    int rc;
#define RUN(expr) if ((rc = (expr)) != MVS_OK) return rc
    RUN(conv3d_fp32(volume, p->w[0], p->shift[0], 1, c0, B, 32, 8, D, H, W, 1, st));
    RUN(conv3d_fp32(c0, p->w[1], p->shift[1], 1, c1, B, 8, 16, D, H, W, 2, st));
    RUN(conv3d_fp32(c1, p->w[2], p->shift[2], 1, c2, B, 16, 16, D / 2, H / 2, W / 2, 1, st));
    RUN(conv3d_fp32(c2, p->w[3], p->shift[3], 1, c3, B, 16, 32, D / 2, H / 2, W / 2, 2, st));
    RUN(conv3d_fp32(c3, p->w[4], p->shift[4], 1, c4, B, 32, 32, D / 4, H / 4, W / 4, 1, st));
    RUN(conv3d_fp32(c4, p->w[5], p->shift[5], 1, c5, B, 32, 64, D / 4, H / 4, W / 4, 2, st));
    RUN(conv3d_fp32(c5, p->w[6], p->shift[6], 1, c6, B, 64, 64, D / 8, H / 8, W / 8, 1, st));
    RUN(convT3d_fp32(c6, p->w[7], p->shift[7], 1, c4, u7, B, 64, 32, D / 8, H / 8, W / 8, st));
    RUN(convT3d_fp32(u7, p->w[8], p->shift[8], 1, c2, u9, B, 32, 16, D / 4, H / 4, W / 4, st));
    RUN(convT3d_fp32(u9, p->w[9], p->shift[9], 1, c0, u11, B, 16, 8, D / 2, H / 2, W / 2, st));
    RUN(conv3d_fp32(u11, p->w[10], p->shift[10], 0, logits, B, 8, 1, D, H, W, 1, st));
#undef RUN
    return MVS_OK;
}

// bf16 chunk-planar ("CP8") variants: the fused warp+variance kernel writes the tensor-core CostRegNet's input
// layout directly, so the fp32 volume never exists in the bf16 precision mode.
extern "C" size_t mvs_volume_cp8_bytes(int B, int D, int H, int W) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)B * 32 * D * H * W * 2;
}

extern "C" int mvs_warp_variance_fwd_cp8(const float *fea, const float *proj, const float *depth_values, void *vol_cp8,
                                         void *workspace, int B, int V, int C, int D, int H, int W, void *stream) {
    MVS_REQUIRE(fea && proj && depth_values && vol_cp8 && workspace, "null pointer argument");
    MVS_REQUIRE(C == 32, "warp_variance: C must be 32, got %d", C);
    MVS_REQUIRE(B > 0 && V >= 1 && V <= 64 && D > 0 && H > 1 && W > 1, "bad shape");
    MVS_REQUIRE((long long)B * D <= 65535LL * 16 && (long long)H * W < (1LL << 27), "shape too large");
    return warp_variance_cp8(fea, proj, depth_values, vol_cp8, workspace, B, V, D, H, W, (cudaStream_t)stream);
}

extern "C" int mvs_warp_variance_fwd_cp8_f16(const void *fea16_nhwc, const float *proj, const float *depth_values,
                                             void *vol_cp8, void *workspace, int B, int V, int C, int D, int H, int W,
                                             void *stream) {
    MVS_REQUIRE(fea16_nhwc && proj && depth_values && vol_cp8 && workspace, "null pointer argument");
    MVS_REQUIRE(C == 32, "warp_variance: C must be 32, got %d", C);
    MVS_REQUIRE(B > 0 && V >= 1 && V <= 64 && D > 0 && H > 1 && W > 1, "bad shape");
    MVS_REQUIRE((long long)B * D <= 65535LL * 16 && (long long)H * W < (1LL << 27), "shape too large");
    MVS_REQUIRE(((uintptr_t)fea16_nhwc & 15) == 0, "fea16_nhwc must be 16-byte aligned");
    return warp_variance_cp8_f16(fea16_nhwc, proj, depth_values, vol_cp8, workspace, B, V, D, H, W, (cudaStream_t)stream);
}

extern "C" int mvs_warp_variance_fwd_cp8_feat(const void *fea_rcp8_f16, const float *proj, const float *depth_values,
                                              void *vol_cp8, void *workspace, int B, int V, int C, int D, int H, int W,
                                              void *stream) {
    MVS_REQUIRE(fea_rcp8_f16 && proj && depth_values && vol_cp8 && workspace, "null pointer argument");
    MVS_REQUIRE(C == 32, "warp_variance: C must be 32, got %d", C);
    MVS_REQUIRE(B > 0 && V >= 1 && V <= 64 && D > 0 && H > 1 && W > 1, "bad shape");
    MVS_REQUIRE((long long)B * D <= 65535LL * 16 && (long long)H * W < (1LL << 27), "shape too large");
    MVS_REQUIRE(((uintptr_t)fea_rcp8_f16 & 15) == 0, "features must be 16-byte aligned");
    return warp_variance_cp8_rcp8(fea_rcp8_f16, proj, depth_values, vol_cp8, workspace, B, V, D, H, W, (cudaStream_t)stream);
}

extern "C" int mvs_warp_variance_fwd_cp8_pool(const void *pool_rcp8_f16, int n_pool, const int *view_ids_host,
                                              const float *proj, const float *depth_values, void *vol_cp8, void *workspace,
                                              int B, int V, int C, int D, int H, int W, void *stream) {
    MVS_REQUIRE(pool_rcp8_f16 && view_ids_host && proj && depth_values && vol_cp8 && workspace, "null pointer argument");
    MVS_REQUIRE(C == 32, "warp_variance: C must be 32, got %d", C);
    MVS_REQUIRE(B > 0 && V >= 1 && V <= 64 && D > 0 && H > 1 && W > 1 && n_pool > 0, "bad shape");
    MVS_REQUIRE((long long)B * D <= 65535LL * 16 && (long long)H * W < (1LL << 27), "shape too large");
    MVS_REQUIRE(((uintptr_t)pool_rcp8_f16 & 15) == 0, "features must be 16-byte aligned");
    return warp_variance_cp8_pool(pool_rcp8_f16, n_pool, view_ids_host, proj, depth_values, vol_cp8, workspace, B, V, D, H, W,
                                  (cudaStream_t)stream);
}

extern "C" int mvs_costreg_fwd_cp8(const void *vol_cp8, const mvs_costreg_params *p, float *logits, void *workspace,
                                   int B, int D, int H, int W, void *stream) {
    MVS_REQUIRE(vol_cp8 && p && logits && workspace, "null pointer argument");
    MVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && D % 8 == 0 && H % 8 == 0 && W % 8 == 0,
                "CostRegNet needs D, H, W divisible by 8 (got D=%d H=%d W=%d)", D, H, W);
    for (int i = 0; i < MVS_COSTREG_LAYERS; ++i)
        MVS_REQUIRE(p->w[i] && p->shift[i], "costreg params: layer %d has a null pointer", i);
    return costreg_tc(nullptr, vol_cp8, p, logits, workspace, B, D, H, W, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Host-buffer entry point (see header).
// ---------------------------------------------------------------------------------------------
namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
};
}  // namespace

extern "C" int mvs_depth_from_features_host(const float *fea_host, const float *proj_host,
                                            const float *depth_values_host, const mvs_costreg_params *params_host,
                                            float *depth_host, float *conf_host, int B, int V, int D, int H, int W,
                                            int precision, int device) {
    MVS_REQUIRE(fea_host && proj_host && depth_values_host && params_host && depth_host && conf_host,
                "null pointer argument");
    MVS_REQUIRE(B > 0 && V >= 1 && D > 0 && H > 1 && W > 1, "bad shape");
    MVS_REQUIRE(D % 8 == 0 && H % 8 == 0 && W % 8 == 0, "D, H, W must be divisible by 8");
    MVS_CUDA(cudaSetDevice(device));
    cudaStream_t st;
    MVS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard {
        cudaStream_t s;
        ~StreamGuard() { cudaStreamDestroy(s); }
    } guard{st};

    const size_t HW = (size_t)H * W, n0 = (size_t)D * HW;
    const size_t fea_b = (size_t)B * V * 32 * HW * 4, proj_b = (size_t)B * V * 64, dv_b = (size_t)B * D * 4;
    // the weights live in device memory owned by this call: packed copies keyed by those pointers must not outlive it, on
    // any exit path (a later cudaMalloc may return the same addresses).  Declared first = destroyed last, after `wts`... the
    // clear only drops cache entries, it does not touch the buffers, so the order against ~DevBuf does not matter.
    struct CacheGuard {
        bool on;
        ~CacheGuard() { if (on) mvs_weight_cache_clear(); }
    } cache_guard{precision == MVS_PRECISION_BF16};
    DevBuf fea, proj, dv, var, ws1, ws2, logits, depth, conf, wts;
    MVS_CUDA(fea.alloc(fea_b));
    MVS_CUDA(proj.alloc(proj_b));
    MVS_CUDA(dv.alloc(dv_b));
    MVS_CUDA(var.alloc((size_t)B * 32 * n0 * 4));
    MVS_CUDA(ws1.alloc(mvs_warp_variance_workspace_bytes(B, V, 32, H, W)));
    MVS_CUDA(ws2.alloc(mvs_costreg_workspace_bytes(B, D, H, W, precision)));
    MVS_CUDA(logits.alloc((size_t)B * n0 * 4));
    MVS_CUDA(depth.alloc((size_t)B * HW * 4));
    MVS_CUDA(conf.alloc((size_t)B * HW * 4));
    size_t wcount = 0;
    for (int i = 0; i < MVS_COSTREG_LAYERS; ++i) wcount += (size_t)kLayerCin[i] * kLayerCout[i] * 27 + kLayerCout[i];
    MVS_CUDA(wts.alloc(wcount * 4));
    mvs_costreg_params dp;
    {
        float *cur = (float *)wts.p;
        for (int i = 0; i < MVS_COSTREG_LAYERS; ++i) {
            const size_t nw = (size_t)kLayerCin[i] * kLayerCout[i] * 27;
            MVS_REQUIRE(params_host->w[i] && params_host->shift[i], "params_host: layer %d has a null pointer", i);
            MVS_CUDA(cudaMemcpyAsync(cur, params_host->w[i], nw * 4, cudaMemcpyHostToDevice, st));
            dp.w[i] = cur;
            cur += nw;
            MVS_CUDA(cudaMemcpyAsync(cur, params_host->shift[i], (size_t)kLayerCout[i] * 4, cudaMemcpyHostToDevice, st));
            dp.shift[i] = cur;
            cur += kLayerCout[i];
        }
    }
    MVS_CUDA(cudaMemcpyAsync(fea.p, fea_host, fea_b, cudaMemcpyHostToDevice, st));
    MVS_CUDA(cudaMemcpyAsync(proj.p, proj_host, proj_b, cudaMemcpyHostToDevice, st));
    MVS_CUDA(cudaMemcpyAsync(dv.p, depth_values_host, dv_b, cudaMemcpyHostToDevice, st));
    int rc;
    if ((rc = mvs_warp_variance_fwd((const float *)fea.p, (const float *)proj.p, (const float *)dv.p, (float *)var.p,
                                    ws1.p, B, V, 32, D, H, W, st)) != MVS_OK)
        return rc;
    if ((rc = mvs_costreg_fwd((const float *)var.p, &dp, (float *)logits.p, ws2.p, B, D, H, W, precision, st)) != MVS_OK)
        return rc;
    if ((rc = mvs_softmax_depth_conf((const float *)logits.p, (const float *)dv.p, (float *)depth.p, (float *)conf.p,
                                     nullptr, B, D, H, W, st)) != MVS_OK)
        return rc;
    MVS_CUDA(cudaMemcpyAsync(depth_host, depth.p, (size_t)B * HW * 4, cudaMemcpyDeviceToHost, st));
    MVS_CUDA(cudaMemcpyAsync(conf_host, conf.p, (size_t)B * HW * 4, cudaMemcpyDeviceToHost, st));
    MVS_CUDA(cudaStreamSynchronize(st));
    return MVS_OK;
}
