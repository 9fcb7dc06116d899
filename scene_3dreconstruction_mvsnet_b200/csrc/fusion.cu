// Geometric-consistency filter of the depth maps -- the step right after the depth-inference path
// (SURVEY.md section 8(f) rank 3).  Replaces, for the reference (olivier-2018/scene_3Dreconstruction_MVSNet):
//   eval.py:508-560  reproject_with_depth         project reference pixels into a source view, sample its depth map
//                                                 (cv2.remap, bilinear), project the sampled points back
//   eval.py:564-585  check_geometric_consistency  |p_reproj - p| < condmask_pixel  and  |d_reproj - d| / d < condmask_depth
//   eval.py:660-703  filter_depth (per reference view): photometric mask, per-source masks, averaged depth,
//                                                 geometric mask (>= geomask consistent views), final mask
// The reference runs ~60 full-image numpy passes and one cv2.remap per (reference, source) pair on the CPU; here one
// thread owns one reference pixel and walks all source views: a depth map pair costs 12 bytes of HBM traffic per pixel
// and view, everything else stays in registers.  Arithmetic follows the reference's types: float64 for the projections
// (numpy promotes), float32 where the reference casts (.astype(np.float32)), and cv2.remap's fixed-point sampling
// (coordinates rounded to 1/32 pixel, float32 weights, left-to-right float32 accumulation, constant border 0).
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"

namespace mvs {

namespace {

// per source view: everything the per-pixel chain needs, composed on the host in double
struct FusionView {
    double T1[12];    // (E_src * inv(E_ref))[:3, :4]      eval.py:529
    double Ksrc[9];   //                                     eval.py:532
    double KsrcInv[9];  //                                   eval.py:545
    double T2[12];    // (E_ref * inv(E_src))[:3, :4]      eval.py:548
};
struct FusionParams {
    double KrefInv[9];  // eval.py:526
    double Kref[9];     // eval.py:552
    double pix_thr;     // condmask_pixel (float64 comparison: dist is float64 in the reference)
    float dep_thr;      // condmask_depth (float32 comparison: relative_depth_diff is float32, NEP 50 weak scalar)
    float photo_thr;    // photomask
    int geomask, S, H, W;
};

__device__ __forceinline__ void mat3(const double *m, double x, double y, double z, double &ox, double &oy, double &oz) {
    ox = __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), __dmul_rn(m[2], z));
    oy = __dadd_rn(__dadd_rn(__dmul_rn(m[3], x), __dmul_rn(m[4], y)), __dmul_rn(m[5], z));
    oz = __dadd_rn(__dadd_rn(__dmul_rn(m[6], x), __dmul_rn(m[7], y)), __dmul_rn(m[8], z));
}
__device__ __forceinline__ void mat34(const double *m, double x, double y, double z, double &ox, double &oy, double &oz) {
    ox = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), __dmul_rn(m[2], z)), m[3]);
    oy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[4], x), __dmul_rn(m[5], y)), __dmul_rn(m[6], z)), m[7]);
    oz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[8], x), __dmul_rn(m[9], y)), __dmul_rn(m[10], z)), m[11]);
}

// cvRound(v) as cv2 computes it on x86 (cvtss2si): nearest even; NaN and out-of-range give INT_MIN
__device__ __forceinline__ int cv_round(float v) {
    if (!(fabsf(v) <= 2147483520.f)) return INT_MIN;
    return __float2int_rn(v);
}

// cv2.remap(src, x, y, INTER_LINEAR) for one sample, float32 image, BORDER_CONSTANT 0
__device__ __forceinline__ float remap_bilinear(const float *__restrict__ src, int H, int W, float x, float y) {
    const int sx = cv_round(__fmul_rn(x, 32.f)), sy = cv_round(__fmul_rn(y, 32.f));
    const float ax = (float)(sx & 31) * 0.03125f, ay = (float)(sy & 31) * 0.03125f;  // exact: k / 32
    const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767);
    const float w00 = __fmul_rn(1.f - ay, 1.f - ax), w01 = __fmul_rn(1.f - ay, ax);
    const float w10 = __fmul_rn(ay, 1.f - ax), w11 = __fmul_rn(ay, ax);
    auto tap = [&](int yy, int xx) -> float {
        return ((unsigned)xx < (unsigned)W && (unsigned)yy < (unsigned)H) ? __ldg(src + (size_t)yy * W + xx) : 0.f;
    };
    float v = __fmul_rn(tap(iy, ix), w00);
    v = __fadd_rn(v, __fmul_rn(tap(iy, ix + 1), w01));
    v = __fadd_rn(v, __fmul_rn(tap(iy + 1, ix), w10));
    v = __fadd_rn(v, __fmul_rn(tap(iy + 1, ix + 1), w11));
    return v;
}

__global__ void __launch_bounds__(128)
filter_depth_kernel(const float *__restrict__ ref_depth, const float *__restrict__ confidence,
                    const float *__restrict__ src_depths, const FusionView *__restrict__ views, const FusionParams P,
                    double *__restrict__ depth_avg, int32_t *__restrict__ geo_sum_out, uint8_t *__restrict__ photo_mask,
                    uint8_t *__restrict__ geo_mask, uint8_t *__restrict__ final_mask, float *__restrict__ reprojected,
                    uint8_t *__restrict__ src_masks, float *__restrict__ xy_src) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.W) return;
    const size_t HW = (size_t)P.H * P.W, pix = (size_t)y * P.W + x;
    const float d = __ldg(ref_depth + pix);
    const double dd = (double)d;
    // xyz_ref = inv(K_ref) @ ((x, y, 1) * depth)                                        eval.py:526
    double rx, ry, rz;
    mat3(P.KrefInv, __dmul_rn((double)x, dd), __dmul_rn((double)y, dd), dd, rx, ry, rz);
    float sum = 0.f;  // python sum() of float32 arrays: 0 + a_0 + a_1 + ... in float32
    int geo_sum = 0;
    for (int s = 0; s < P.S; ++s) {
        const FusionView &V = views[s];
        double sx, sy, sz, kx, ky, kz;
        mat34(V.T1, rx, ry, rz, sx, sy, sz);                                          // eval.py:529
        mat3(V.Ksrc, sx, sy, sz, kx, ky, kz);                                         // eval.py:532
        const double xs = __ddiv_rn(kx, kz), ys = __ddiv_rn(ky, kz);                  // eval.py:533
        const float x_src = (float)xs, y_src = (float)ys;                             // eval.py:539-540
        const float sampled = remap_bilinear(src_depths + (size_t)s * HW, P.H, P.W, x_src, y_src);  // eval.py:541
        const double sd = (double)sampled;
        double bx, by, bz, px, py, pz;
        mat3(V.KsrcInv, __dmul_rn(xs, sd), __dmul_rn(ys, sd), sd, bx, by, bz);        // eval.py:546
        mat34(V.T2, bx, by, bz, px, py, pz);                                          // eval.py:549
        const float depth_rep = (float)pz;                                            // eval.py:552
        mat3(P.Kref, px, py, pz, kx, ky, kz);                                         // eval.py:553
        const float x_rep = (float)__ddiv_rn(kx, kz), y_rep = (float)__ddiv_rn(ky, kz);  // eval.py:554-556
        // |p_reproj - p| (float64: float32 array minus int64 grid), |d_reproj - d| / d (float32)    eval.py:572-577
        const double ex = __dsub_rn((double)x_rep, (double)x), ey = __dsub_rn((double)y_rep, (double)y);
        const double dist = sqrt(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
        const float rel = __fdiv_rn(fabsf(__fsub_rn(depth_rep, d)), d);
        const bool m = (dist < P.pix_thr) && (rel < P.dep_thr);                       // eval.py:580
        const float kept = m ? depth_rep : 0.f;                                       // eval.py:583
        sum = __fadd_rn(sum, kept);
        geo_sum += m ? 1 : 0;
        if (reprojected) reprojected[(size_t)s * HW + pix] = kept;
        if (src_masks) src_masks[(size_t)s * HW + pix] = m ? 1 : 0;
        if (xy_src) {
            xy_src[((size_t)s * 2) * HW + pix] = x_src;
            xy_src[((size_t)s * 2 + 1) * HW + pix] = y_src;
        }
    }
    // depth_est_averaged = (sum(reprojected) + ref_depth) / (geo_mask_sum + 1): float32 numerator, float64 quotient    eval.py:700
    depth_avg[pix] = __ddiv_rn((double)__fadd_rn(sum, d), (double)(geo_sum + 1));
    geo_sum_out[pix] = geo_sum;
    const bool pm = confidence ? (__ldg(confidence + pix) > P.photo_thr) : true;      // eval.py:660
    const bool gm = geo_sum >= P.geomask;                                             // eval.py:702
    photo_mask[pix] = pm ? 1 : 0;
    geo_mask[pix] = gm ? 1 : 0;
    final_mask[pix] = (pm && gm) ? 1 : 0;                                             // eval.py:703
}

// ---- small host-side linear algebra (double), for the camera compositions numpy does with LAPACK ----
bool invert_n(const double *m, double *inv, int n) {
    double a[16], b[16];
    for (int i = 0; i < n * n; ++i) { a[i] = m[i]; b[i] = 0; }
    for (int i = 0; i < n; ++i) b[i * n + i] = 1;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (fabs(a[r * n + c]) > fabs(a[piv * n + c])) piv = r;
        if (a[piv * n + c] == 0.0) return false;
        if (piv != c)
            for (int k = 0; k < n; ++k) { std::swap(a[c * n + k], a[piv * n + k]); std::swap(b[c * n + k], b[piv * n + k]); }
        const double p = a[c * n + c];
        for (int k = 0; k < n; ++k) { a[c * n + k] /= p; b[c * n + k] /= p; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[r * n + c];
            if (f == 0.0) continue;
            for (int k = 0; k < n; ++k) { a[r * n + k] -= f * a[c * n + k]; b[r * n + k] -= f * b[c * n + k]; }
        }
    }
    for (int i = 0; i < n * n; ++i) inv[i] = b[i];
    return true;
}
void mul44(const double *a, const double *b, double *o) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            o[i * 4 + j] = s;
        }
}

}  // namespace
}  // namespace mvs

using namespace mvs;

extern "C" int mvs_filter_depth(const float *ref_depth, const float *confidence, const double *ref_K_host,
                                const double *ref_E_host, const float *src_depths, const double *src_K_host,
                                const double *src_E_host, int S, int H, int W, double condmask_pixel, double condmask_depth,
                                int geomask, double photomask, double *depth_avg, int32_t *geo_mask_sum, uint8_t *photo_mask,
                                uint8_t *geo_mask, uint8_t *final_mask, float *reprojected, uint8_t *src_masks, float *xy_src,
                                void *stream) {
    MVS_REQUIRE(ref_depth && ref_K_host && ref_E_host && depth_avg && geo_mask_sum && photo_mask && geo_mask && final_mask,
                "null pointer argument");
    MVS_REQUIRE(S >= 0 && S <= 64 && H > 0 && W > 0, "bad shape S=%d H=%d W=%d", S, H, W);
    MVS_REQUIRE(S == 0 || (src_depths && src_K_host && src_E_host), "null source-view argument");
    cudaStream_t st = (cudaStream_t)stream;
    FusionParams P;
    double ref_E_inv[16];
    MVS_REQUIRE(invert_n(ref_K_host, P.KrefInv, 3), "reference intrinsics are singular");
    MVS_REQUIRE(invert_n(ref_E_host, ref_E_inv, 4), "reference extrinsics are singular");
    for (int i = 0; i < 9; ++i) P.Kref[i] = ref_K_host[i];
    P.pix_thr = condmask_pixel;
    P.dep_thr = (float)condmask_depth;
    P.photo_thr = (float)photomask;
    P.geomask = geomask; P.S = S; P.H = H; P.W = W;
    FusionView hv[64];
    for (int s = 0; s < S; ++s) {
        const double *K = src_K_host + 9 * s, *E = src_E_host + 16 * s;
        double Einv[16], t1[16], t2[16];
        MVS_REQUIRE(invert_n(K, hv[s].KsrcInv, 3), "source view %d: intrinsics are singular", s);
        MVS_REQUIRE(invert_n(E, Einv, 4), "source view %d: extrinsics are singular", s);
        mul44(E, ref_E_inv, t1);
        mul44(ref_E_host, Einv, t2);
        for (int i = 0; i < 12; ++i) { hv[s].T1[i] = t1[i]; hv[s].T2[i] = t2[i]; }
        for (int i = 0; i < 9; ++i) hv[s].Ksrc[i] = K[i];
    }
    FusionView *dv = nullptr;
    if (S > 0) {
        MVS_CUDA(cudaMallocAsync((void **)&dv, sizeof(FusionView) * S, st));
        MVS_CUDA(cudaMemcpyAsync(dv, hv, sizeof(FusionView) * S, cudaMemcpyHostToDevice, st));  // pageable source: staged before return
    }
    dim3 grid(cdiv(W, 128), H);
    filter_depth_kernel<<<grid, 128, 0, st>>>(ref_depth, confidence, src_depths, dv, P, depth_avg, geo_mask_sum, photo_mask, geo_mask,
                                             final_mask, reprojected, src_masks, xy_src);
    cudaError_t e = cudaGetLastError();
    if (dv) cudaFreeAsync(dv, st);
    if (e != cudaSuccess) return set_error(MVS_ERR_CUDA, "filter_depth launch failed: %s", cudaGetErrorString(e));
    count_launches(1);
    return MVS_OK;
}
