// Plane-sweep homography warp fused with the variance cost-volume reduction (forward + backward).
//
// Replaces, for the reference (olivier-2018/scene_3Dreconstruction_MVSNet):
//   models/module.py:96-139   homo_warping         (grid construction + F.grid_sample)
//   models/mvsnet.py:145-177  running sum / sum of squares over views and the variance
//   autograd through both     (backward kernels at the bottom of this file)
//
// Design (B200 / sm_100a, HBM-bound output, L1/LSU-bound input side):
//   * source-view features are first re-laid channels-last (NHWC, 32 ch = one 128-byte line per
//     texel) so that one bilinear tap of all 32 channels is exactly one L1 line;
//   * a warp owns 32 consecutive x of one (b, y) row and a chunk of depth planes.  Lane L computes
//     the sample position of pixel L once per (view, depth) -- the reference's exact fp32 operation
//     order, no FMA contraction -- and publishes 4 clamped texel offsets + 4 zero-masked weights
//     through shared memory;
//   * for the gather, lanes are re-mapped as (p = L/8, g = L%8): the 8 lanes of a group read the
//     8 float4 of ONE texel line (a full 128-byte wavefront), the 4 groups process 4 pixels per
//     step.  Thread (p, g) owns pixels x0+8p..x0+8p+7 and channels 4g..4g+3, so the running
//     sum / sum-of-squares live in 64 registers and every store is a 128-bit store with the 4
//     groups of a warp covering one full 128-byte line of the NCDHW volume;
//   * per-view warped volumes never exist in memory: HBM traffic is the algorithmic
//     4*B*32*D*H*W bytes written + the feature maps read once.
#include <algorithm>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"

namespace mvs {

constexpr int kC = 32;          // feature channels (FeatureNet width, mvsnet.py:24)
constexpr int kWarps = 8;       // rows per CTA
constexpr int kThreads = kWarps * 32;

// ------------------------------------------------------------------------------------------------
// proj = src_proj @ inverse(ref_proj)  (module.py:107); rot = proj[:3,:3], trans = proj[:3,3].
// Computed in double with explicitly rounded operations (no contraction) and rounded once to fp32:
// bit-identical to oracle/mvsnet_oracle.c:orc_compose_homography.
// ------------------------------------------------------------------------------------------------
__device__ bool invert4x4(const double *m, double *inv) {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            a[i][j] = m[i * 4 + j];
            a[i][j + 4] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r)
            if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if (a[p][c] == 0.0) return false;
        if (p != c)
            for (int j = 0; j < 8; ++j) {
                double t = a[c][j];
                a[c][j] = a[p][j];
                a[p][j] = t;
            }
        double d = __ddiv_rn(1.0, a[c][c]);
        for (int j = 0; j < 8; ++j) a[c][j] = __dmul_rn(a[c][j], d);
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            double f = a[r][c];
            if (f != 0.0)
                for (int j = 0; j < 8; ++j) a[r][j] = __dsub_rn(a[r][j], __dmul_rn(f, a[c][j]));
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) inv[i * 4 + j] = a[i][j + 4];
    return true;
}

__device__ void compose_one(const float *src_proj, const float *ref_proj, float *rt) {
    double s[16], r[16], ri[16];
    for (int i = 0; i < 16; ++i) {
        s[i] = (double)src_proj[i];
        r[i] = (double)ref_proj[i];
    }
    if (!invert4x4(r, ri)) {  // singular reference camera: poison the output like torch.inverse would fail
        for (int i = 0; i < 12; ++i) rt[i] = __int_as_float(0x7fc00000);
        return;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc = __dadd_rn(acc, __dmul_rn(s[i * 4 + k], ri[k * 4 + j]));
            if (j < 3) rt[i * 3 + j] = (float)acc;
            else rt[9 + i] = (float)acc;
        }
}

__global__ void compose_views_kernel(const float *__restrict__ proj, float *__restrict__ rt, int B, int V) {
    // programmatic dependent launch: the warp kernel behind may be scheduled early; nothing is read or written before
    // every earlier kernel of the stream has completed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B * (V - 1)) return;
    int b = n / (V - 1), v = n % (V - 1) + 1;
    compose_one(proj + ((size_t)b * V + v) * 16, proj + (size_t)b * V * 16, rt + (size_t)n * 12);
}

__global__ void compose_pairs_kernel(const float *__restrict__ src_proj, const float *__restrict__ ref_proj,
                                     float *__restrict__ rt, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    compose_one(src_proj + (size_t)b * 16, ref_proj + (size_t)b * 16, rt + (size_t)b * 12);
}

int compose_homographies(const float *proj, float *rt, int B, int V, cudaStream_t st) {
    int n = B * (V - 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cdiv(n, 64));
    cfg.blockDim = dim3(64);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MVS_CUDA(cudaLaunchKernelEx(&cfg, compose_views_kernel, proj, (float *)rt, B, V));
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

int compose_homography_pairs(const float *src_proj, const float *ref_proj, float *rt, int B, cudaStream_t st) {
    compose_pairs_kernel<<<cdiv(B, 64), 64, 0, st>>>(src_proj, ref_proj, rt, B);
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

// ------------------------------------------------------------------------------------------------
// Layout changes: [N][32][HW] <-> [N][HW][32].  `view_stride`/`views_per_batch` let the forward
// kernel pick the V-1 source views out of a [B,V,C,H,W] tensor without a gather copy.
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc32_kernel(const float *__restrict__ in, float *__restrict__ out, int HW, int nsrc, int V) {
    __shared__ float tile[32][33];
    const int n = blockIdx.y;  // index over B*(nsrc)
    const int b = n / nsrc, v = n % nsrc + (V - nsrc);
    const float *src = in + ((size_t)b * V + v) * kC * HW;
    float *dst = out + (size_t)n * HW * kC;
    const int p0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int c = ty; c < 32; c += 8) tile[c][tx] = (p0 + tx < HW) ? src[(size_t)c * HW + p0 + tx] : 0.f;
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (p0 + r < HW) dst[(size_t)(p0 + r) * kC + tx] = tile[tx][r];
}

// grads: [N][HW][32] -> rows of a [B,V,C,H,W] tensor (views V-nsrc..V-1), overwrite
__global__ void nhwc32_to_nchw_kernel(const float *__restrict__ in, float *__restrict__ out, int HW, int nsrc, int V) {
    __shared__ float tile[32][33];
    const int n = blockIdx.y;
    const int b = n / nsrc, v = n % nsrc + (V - nsrc);
    const float *src = in + (size_t)n * HW * kC;
    float *dst = out + ((size_t)b * V + v) * kC * HW;
    const int p0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += 8) tile[r][tx] = (p0 + r < HW) ? src[(size_t)(p0 + r) * kC + tx] : 0.f;
    __syncthreads();
    for (int c = ty; c < 32; c += 8)
        if (p0 + tx < HW) dst[(size_t)c * HW + p0 + tx] = tile[tx][c];
}

// ------------------------------------------------------------------------------------------------
// Sample position of module.py:119-136 for pixel (x, y) at depth d, reference operation order.
// Produces 4 clamped texel indices (y*W + x) and 4 weights, zeroed where the tap is outside the
// image (grid_sample padding_mode='zeros') or the coordinate is not finite (CUDA grid_sampler rule).
// ------------------------------------------------------------------------------------------------
struct Taps {
    int4 off;
    float4 w;  // nw, ne, sw, se
};

__device__ __forceinline__ float safe_coord(float v) {
    return (v <= 2147483520.0f && v >= -2147483648.0f) ? v : -100.0f;  // NaN fails both compares
}

__device__ __forceinline__ Taps sample_taps(const float *__restrict__ rt, float x, float y, float d, int H, int W) {
    const float rx = __fadd_rn(__fadd_rn(__fmul_rn(rt[0], x), __fmul_rn(rt[1], y)), rt[2]);
    const float ry = __fadd_rn(__fadd_rn(__fmul_rn(rt[3], x), __fmul_rn(rt[4], y)), rt[5]);
    const float rz = __fadd_rn(__fadd_rn(__fmul_rn(rt[6], x), __fmul_rn(rt[7], y)), rt[8]);
    const float qx = __fadd_rn(__fmul_rn(rx, d), rt[9]);
    const float qy = __fadd_rn(__fmul_rn(ry, d), rt[10]);
    const float qz = __fadd_rn(__fmul_rn(rz, d), rt[11]);
    const float px = __fdiv_rn(qx, qz);
    const float py = __fdiv_rn(qy, qz);
    const float gx = __fsub_rn(__fdiv_rn(px, (float)(W - 1) * 0.5f), 1.0f);
    const float gy = __fsub_rn(__fdiv_rn(py, (float)(H - 1) * 0.5f), 1.0f);
    const float ix = safe_coord(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)W), 1.0f), 0.5f));
    const float iy = safe_coord(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)H), 1.0f), 0.5f));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const float ax = __fsub_rn(__fadd_rn(fx0, 1.0f), ix), bx = __fsub_rn(ix, fx0);
    const float ay = __fsub_rn(__fadd_rn(fy0, 1.0f), iy), by = __fsub_rn(iy, fy0);
    const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
    const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
    const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
    const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
    Taps t;
    t.off = make_int4(cy0 * W + cx0, cy0 * W + cx1, cy1 * W + cx0, cy1 * W + cx1);
    t.w.x = (vx0 && vy0) ? __fmul_rn(ax, ay) : 0.f;
    t.w.y = (vx1 && vy0) ? __fmul_rn(bx, ay) : 0.f;
    t.w.z = (vx0 && vy1) ? __fmul_rn(ax, by) : 0.f;
    t.w.w = (vx1 && vy1) ? __fmul_rn(bx, by) : 0.f;
    return t;
}

__device__ __forceinline__ float4 bilerp4(const float4 *__restrict__ f, const int4 o, const float4 w) {
    const float4 a = __ldg(f + (size_t)o.x * 8), b = __ldg(f + (size_t)o.y * 8);
    const float4 c = __ldg(f + (size_t)o.z * 8), d = __ldg(f + (size_t)o.w * 8);
    float4 r;
    r.x = fmaf(d.x, w.w, fmaf(c.x, w.z, fmaf(b.x, w.y, a.x * w.x)));
    r.y = fmaf(d.y, w.w, fmaf(c.y, w.z, fmaf(b.y, w.y, a.y * w.x)));
    r.z = fmaf(d.z, w.w, fmaf(c.z, w.z, fmaf(b.z, w.y, a.z * w.x)));
    r.w = fmaf(d.w, w.w, fmaf(c.w, w.z, fmaf(b.w, w.y, a.w * w.x)));
    return r;
}

// ------------------------------------------------------------------------------------------------
// Forward.  MODE_VAR: out = variance over (ref + nsrc source views).  MODE_WARP: nsrc == 1,
// out = the warped source volume itself (the standalone, materialising homo_warping).
// grid = (ceil(W/32), ceil(H/8), B * ceil(D/dchunk)), block = 256.
// ------------------------------------------------------------------------------------------------
enum { MODE_VAR = 0, MODE_WARP = 1 };

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
warp_volume_fwd_kernel(const float *__restrict__ fea,       // [B,V,32,H,W]; view 0 = ref (MODE_VAR only)
                       const float4 *__restrict__ src_cl,    // [B*nsrc][H*W][8] float4, channels-last
                       const float *__restrict__ rt,         // [B*nsrc][12]
                       const float *__restrict__ depth_values,  // [B,D]
                       float *__restrict__ out,              // [B,32,D,H,W]
                       int V, int nsrc, int D, int H, int W, int dchunk) {
    __shared__ float4 s_w[kWarps][32];
    __shared__ int4 s_o[kWarps][32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = lane >> 3, g = lane & 7;
    const int nchunks = (D + dchunk - 1) / dchunk;
    const int b = blockIdx.z / nchunks;
    const int d_begin = (blockIdx.z % nchunks) * dchunk;
    const int d_end = min(D, d_begin + dchunk);
    const int y = blockIdx.y * kWarps + warp;
    if (y >= H) return;  // warp-uniform; only __syncwarp below
    const int x0 = blockIdx.x * 32;
    const int xr = x0 + 8 * p;  // first pixel of this thread's 8-pixel run
    const size_t HW = (size_t)H * W;
    const bool vec_ok = ((W & 3) == 0);
    const float invV = 1.0f / (float)V;
    const float xl = (float)(x0 + lane), yf = (float)y;

    for (int d = d_begin; d < d_end; ++d) {
        const float dep = __ldg(depth_values + (size_t)b * D + d);
        float S[8][4], Q[8][4];
        if (MODE == MODE_VAR) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float *r = fea + (((size_t)b * V) * kC + 4 * g + j) * HW + (size_t)y * W + xr;
                float v[8];
                if (vec_ok && xr + 7 < W) {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(r));
                    const float4 c = __ldg(reinterpret_cast<const float4 *>(r) + 1);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                    v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = (xr + i < W) ? __ldg(r + i) : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    S[i][j] = v[i];
                    Q[i][j] = v[i] * v[i];
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) S[i][j] = 0.f;
        }

        for (int v = 0; v < nsrc; ++v) {
            const int n = b * nsrc + v;
            const Taps t = sample_taps(rt + (size_t)n * 12, xl, yf, dep, H, W);
            __syncwarp();  // previous view's readers are done with the slots
            s_w[warp][lane] = t.w;
            s_o[warp][lane] = t.off;
            __syncwarp();
            const float4 *f = src_cl + (size_t)n * HW * 8 + g;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 w = s_w[warp][8 * p + i];
                const int4 o = s_o[warp][8 * p + i];
                const float4 val = bilerp4(f, o, w);
                if (MODE == MODE_VAR) {
                    S[i][0] += val.x; Q[i][0] = fmaf(val.x, val.x, Q[i][0]);
                    S[i][1] += val.y; Q[i][1] = fmaf(val.y, val.y, Q[i][1]);
                    S[i][2] += val.z; Q[i][2] = fmaf(val.z, val.z, Q[i][2]);
                    S[i][3] += val.w; Q[i][3] = fmaf(val.w, val.w, Q[i][3]);
                } else {
                    S[i][0] = val.x; S[i][1] = val.y; S[i][2] = val.z; S[i][3] = val.w;
                }
            }
        }

#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == MODE_VAR) {
                    const float m = S[i][j] * invV;
                    r[i] = fmaf(Q[i][j], invV, -m * m);  // Q/V - (S/V)^2   (mvsnet.py:177)
                } else {
                    r[i] = S[i][j];
                }
            }
            float *o = out + (((size_t)b * kC + 4 * g + j) * D + d) * HW + (size_t)y * W + xr;
            if (vec_ok && xr + 7 < W) {
                __stcs(reinterpret_cast<float4 *>(o), make_float4(r[0], r[1], r[2], r[3]));
                __stcs(reinterpret_cast<float4 *>(o) + 1, make_float4(r[4], r[5], r[6], r[7]));
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (xr + i < W) o[i] = r[i];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Forward, second generation (variance only).  ncu on the first version showed the L1 data pipe
// (l1tex__data_pipe_lsu_wavefronts) at 73 % of peak: every byte delivered to a register costs a
// wavefront slot, so the per-step exchange of 4 offsets + 4 weights (2 x LDS.128 = 8 wavefronts
// against 16 for the taps themselves) was a third of the traffic.  Here lane L publishes the four
// zero-masked interpolation FACTORS (ax, bx, ay, by) and ONE packed base offset (clamped texel
// index + the two clamp-aware increments in bits 30/31): LDS.128 + LDS.32 = 5 wavefronts; the four
// weights are re-formed with 4 multiplies per step.  Masking the factors instead of the products
// gives bit-identical weights (a product with a zeroed factor is the zero the mask would write).
// fp32 texels, the reference's operation order, fp32 NCDHW volume.  Since round 2 the strict-fp32 mode runs the
// TMA-window form of the same arithmetic (warp_variance_win32_kernel in warp_variance_win.cu: 2.46 ms against 4.11 ms at the
// DTU shape, bit-identical volumes); this kernel serves more than 8 source views, where no view gets a window.
// ------------------------------------------------------------------------------------------------
struct PackedTap {
    float4 f;       // ax, bx, ay, by  (zeroed where the corresponding column / row is out of range)
    uint32_t base;  // (cy0*W + cx0) | (cx1 != cx0) << 30 | (cy1 != cy0) << 31
};

__device__ __forceinline__ PackedTap sample_packed(const float *__restrict__ rt, float x, float y, float d, int H, int W) {
    const float rx = __fadd_rn(__fadd_rn(__fmul_rn(rt[0], x), __fmul_rn(rt[1], y)), rt[2]);
    const float ry = __fadd_rn(__fadd_rn(__fmul_rn(rt[3], x), __fmul_rn(rt[4], y)), rt[5]);
    const float rz = __fadd_rn(__fadd_rn(__fmul_rn(rt[6], x), __fmul_rn(rt[7], y)), rt[8]);
    const float qx = __fadd_rn(__fmul_rn(rx, d), rt[9]);
    const float qy = __fadd_rn(__fmul_rn(ry, d), rt[10]);
    const float qz = __fadd_rn(__fmul_rn(rz, d), rt[11]);
    const float px = __fdiv_rn(qx, qz);
    const float py = __fdiv_rn(qy, qz);
    const float gx = __fsub_rn(__fdiv_rn(px, (float)(W - 1) * 0.5f), 1.0f);
    const float gy = __fsub_rn(__fdiv_rn(py, (float)(H - 1) * 0.5f), 1.0f);
    const float ix = safe_coord(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)W), 1.0f), 0.5f));
    const float iy = safe_coord(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)H), 1.0f), 0.5f));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
    const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
    const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
    const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
    PackedTap t;
    t.f.x = vx0 ? __fsub_rn(__fadd_rn(fx0, 1.0f), ix) : 0.f;
    t.f.y = vx1 ? __fsub_rn(ix, fx0) : 0.f;
    t.f.z = vy0 ? __fsub_rn(__fadd_rn(fy0, 1.0f), iy) : 0.f;
    t.f.w = vy1 ? __fsub_rn(iy, fy0) : 0.f;
    t.base = (uint32_t)(cy0 * W + cx0) | ((uint32_t)(cx1 != cx0) << 30) | ((uint32_t)(cy1 != cy0) << 31);
    return t;
}

// Same arithmetic, split in two: R*(x,y,1) depends only on (pixel, view) and is hoisted out of the depth walk.
__device__ __forceinline__ void rot_pixel(const float *__restrict__ rt, float x, float y, float &rx, float &ry, float &rz) {
    rx = __fadd_rn(__fadd_rn(__fmul_rn(rt[0], x), __fmul_rn(rt[1], y)), rt[2]);
    ry = __fadd_rn(__fadd_rn(__fmul_rn(rt[3], x), __fmul_rn(rt[4], y)), rt[5]);
    rz = __fadd_rn(__fadd_rn(__fmul_rn(rt[6], x), __fmul_rn(rt[7], y)), rt[8]);
}

__device__ __forceinline__ PackedTap sample_packed_r(float rx, float ry, float rz, float tx, float ty, float tz, float d,
                                                     int H, int W) {
    const float qx = __fadd_rn(__fmul_rn(rx, d), tx);
    const float qy = __fadd_rn(__fmul_rn(ry, d), ty);
    const float qz = __fadd_rn(__fmul_rn(rz, d), tz);
    const float px = __fdiv_rn(qx, qz);
    const float py = __fdiv_rn(qy, qz);
    const float gx = __fsub_rn(__fdiv_rn(px, (float)(W - 1) * 0.5f), 1.0f);
    const float gy = __fsub_rn(__fdiv_rn(py, (float)(H - 1) * 0.5f), 1.0f);
    const float ix = safe_coord(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)W), 1.0f), 0.5f));
    const float iy = safe_coord(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)H), 1.0f), 0.5f));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
    const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
    const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
    const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
    PackedTap t;
    t.f.x = vx0 ? __fsub_rn(__fadd_rn(fx0, 1.0f), ix) : 0.f;
    t.f.y = vx1 ? __fsub_rn(ix, fx0) : 0.f;
    t.f.z = vy0 ? __fsub_rn(__fadd_rn(fy0, 1.0f), iy) : 0.f;
    t.f.w = vy1 ? __fsub_rn(iy, fy0) : 0.f;
    t.base = (uint32_t)(cy0 * W + cx0) | ((uint32_t)(cx1 != cx0) << 30) | ((uint32_t)(cy1 != cy0) << 31);
    return t;
}

constexpr int kMaxSrcSmem = 8;  // source views whose per-pixel rotation terms are kept in shared memory

__global__ void __launch_bounds__(kThreads, 2)
warp_variance_fwd2_kernel(const float *__restrict__ fea,        // [B,V,32,H,W]; view 0 = reference view
                          const float4 *__restrict__ src_cl,     // [B*nsrc][H*W][8] float4, channels-last
                          const float *__restrict__ rt,          // [B*nsrc][12]
                          const float *__restrict__ depth_values,  // [B,D]
                          float *__restrict__ out,               // [B,32,D,H,W] fp32
                          int V, int nsrc, int D, int H, int W, int dchunk) {
    // Work mapping: a warp owns one 32-pixel row segment and walks `dchunk` consecutive depth planes (measured
    // best among row-/plane-major variants, tools/warp_tune.py).  Everything that does not depend on the plane is
    // hoisted out of the walk: R*(x,y,1) per (pixel, view) and the translation live in shared memory, so the
    // per-(plane, view) coordinate pass issues 6 conflict-free LDS instead of 12 uniform global loads.
    __shared__ float4 s_f[kWarps][32];
    __shared__ uint32_t s_b[kWarps][32];
    __shared__ float s_r[kWarps][kMaxSrcSmem][3][32];
    __shared__ float s_t[kMaxSrcSmem][4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = lane >> 3, g = lane & 7;
    const int nchunks = (D + dchunk - 1) / dchunk;
    const int b = blockIdx.z / nchunks;
    const int d_begin = (blockIdx.z % nchunks) * dchunk;
    const int d_end = min(D, d_begin + dchunk);
    const int y = blockIdx.y * kWarps + warp;
    const int x0 = blockIdx.x * 32;
    const int xr = x0 + 8 * p;
    const size_t HW = (size_t)H * W;
    const bool vec_ok = ((W & 3) == 0);
    const float invV = 1.0f / (float)V;
    const float xl = (float)(x0 + lane), yf = (float)y;
    const int W8 = W * 8;  // float4 units per texel row
    const int nsm = min(nsrc, kMaxSrcSmem);
    if (threadIdx.x < nsm * 3) s_t[threadIdx.x / 3][threadIdx.x % 3] = rt[(size_t)(b * nsrc + threadIdx.x / 3) * 12 + 9 + threadIdx.x % 3];
    for (int v = 0; v < nsm; ++v) {
        float rx, ry, rz;
        rot_pixel(rt + (size_t)(b * nsrc + v) * 12, xl, yf, rx, ry, rz);
        s_r[warp][v][0][lane] = rx;
        s_r[warp][v][1][lane] = ry;
        s_r[warp][v][2][lane] = rz;
    }
    __syncthreads();
    if (y >= H) return;  // warp-uniform; only __syncwarp / full-mask shuffles below

    for (int d = d_begin; d < d_end; ++d) {
        const float dep = __ldg(depth_values + (size_t)b * D + d);
        float S[8][4], Q[8][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float *r = fea + (((size_t)b * V) * kC + 4 * g + j) * HW + (size_t)y * W + xr;
            float v[8];
            if (vec_ok && xr + 7 < W) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(r));
                const float4 c = __ldg(reinterpret_cast<const float4 *>(r) + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (xr + i < W) ? __ldg(r + i) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                S[i][j] = v[i];
                Q[i][j] = v[i] * v[i];
            }
        }

        for (int v = 0; v < nsrc; ++v) {
            const int n = b * nsrc + v;
            const PackedTap t = (v < kMaxSrcSmem)
                                    ? sample_packed_r(s_r[warp][v][0][lane], s_r[warp][v][1][lane], s_r[warp][v][2][lane],
                                                      s_t[v][0], s_t[v][1], s_t[v][2], dep, H, W)
                                    : sample_packed(rt + (size_t)n * 12, xl, yf, dep, H, W);
            __syncwarp();
            s_f[warp][lane] = t.f;
            s_b[warp][lane] = t.base;
            __syncwarp();
            const float4 *f = src_cl + (size_t)n * HW * 8 + g;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 fc = s_f[warp][8 * p + i];
                const uint32_t bb = s_b[warp][8 * p + i];
                const float4 *p00 = f + (size_t)(bb & 0x3FFFFFFFu) * 8;
                const float4 *p01 = p00 + ((bb >> 30) & 1u) * 8;
                const int dyo = (bb >> 31) ? W8 : 0;
                const float4 a = __ldg(p00), bq = __ldg(p01), c = __ldg(p00 + dyo), dq = __ldg(p01 + dyo);
                const float w00 = fc.x * fc.z, w01 = fc.y * fc.z, w10 = fc.x * fc.w, w11 = fc.y * fc.w;
                const float vx = fmaf(dq.x, w11, fmaf(c.x, w10, fmaf(bq.x, w01, a.x * w00)));
                const float vy = fmaf(dq.y, w11, fmaf(c.y, w10, fmaf(bq.y, w01, a.y * w00)));
                const float vz = fmaf(dq.z, w11, fmaf(c.z, w10, fmaf(bq.z, w01, a.z * w00)));
                const float vw = fmaf(dq.w, w11, fmaf(c.w, w10, fmaf(bq.w, w01, a.w * w00)));
                S[i][0] += vx; Q[i][0] = fmaf(vx, vx, Q[i][0]);
                S[i][1] += vy; Q[i][1] = fmaf(vy, vy, Q[i][1]);
                S[i][2] += vz; Q[i][2] = fmaf(vz, vz, Q[i][2]);
                S[i][3] += vw; Q[i][3] = fmaf(vw, vw, Q[i][3]);
            }
        }

        {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float m = S[i][j] * invV;
                    r[i] = fmaf(Q[i][j], invV, -m * m);  // Q/V - (S/V)^2   (mvsnet.py:177)
                }
                float *o = out + (((size_t)b * kC + 4 * g + j) * D + d) * HW + (size_t)y * W + xr;
                if (vec_ok && xr + 7 < W) {
                    __stcs(reinterpret_cast<float4 *>(o), make_float4(r[0], r[1], r[2], r[3]));
                    __stcs(reinterpret_cast<float4 *>(o) + 1, make_float4(r[4], r[5], r[6], r[7]));
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (xr + i < W) o[i] = r[i];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Generic-C fallback of the standalone homo_warping (any channel count, NCHW gathers).  Used only
// when C != 32; one thread per (x, y, d), looping over channels.
__global__ void homo_warp_generic_kernel(const float *__restrict__ src, const float *__restrict__ rt,
                                         const float *__restrict__ depth_values, float *__restrict__ out, int C, int D,
                                         int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int b = blockIdx.z / D, d = blockIdx.z % D;
    if (x >= W) return;
    const size_t HW = (size_t)H * W;
    const Taps t = sample_taps(rt + (size_t)b * 12, (float)x, (float)y, depth_values[(size_t)b * D + d], H, W);
    for (int c = 0; c < C; ++c) {
        const float *f = src + ((size_t)b * C + c) * HW;
        float v = __ldg(f + t.off.x) * t.w.x;
        v = fmaf(__ldg(f + t.off.y), t.w.y, v);
        v = fmaf(__ldg(f + t.off.z), t.w.z, v);
        v = fmaf(__ldg(f + t.off.w), t.w.w, v);
        out[(((size_t)b * C + c) * D + d) * HW + (size_t)y * W + x] = v;
    }
}

__global__ void homo_warp_generic_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ rt,
                                             const float *__restrict__ depth_values, float *__restrict__ gsrc, int C,
                                             int D, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int b = blockIdx.z / D, d = blockIdx.z % D;
    if (x >= W) return;
    const size_t HW = (size_t)H * W;
    const Taps t = sample_taps(rt + (size_t)b * 12, (float)x, (float)y, depth_values[(size_t)b * D + d], H, W);
    for (int c = 0; c < C; ++c) {
        const float gv = gout[(((size_t)b * C + c) * D + d) * HW + (size_t)y * W + x];
        float *f = gsrc + ((size_t)b * C + c) * HW;
        if (t.w.x != 0.f) atomicAdd(f + t.off.x, gv * t.w.x);
        if (t.w.y != 0.f) atomicAdd(f + t.off.y, gv * t.w.y);
        if (t.w.z != 0.f) atomicAdd(f + t.off.z, gv * t.w.z);
        if (t.w.w != 0.f) atomicAdd(f + t.off.w, gv * t.w.w);
    }
}

// ------------------------------------------------------------------------------------------------
// Backward.  Same thread mapping as the forward.  MODE_VAR:
//   mean = S/V,  dL/dx_v = (2/V) (x_v - mean) g  for every view (mvsnet.py:167-177 differentiated);
//   the reference-view gradient is reduced over this CTA's depth chunk in registers and added to
//   grad_fea[:,0] with one atomic per element per chunk; source-view gradients are scattered
//   through the 4 bilinear taps with 128-bit reductions (red.global.add.v4.f32) into a
//   channels-last scratch which is transposed back afterwards.  Warped values are recomputed,
//   never stored (the reference keeps every warped volume alive for autograd).
// MODE_WARP: grad of the standalone homo_warping: scatter g itself.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
warp_volume_bwd_kernel(const float *__restrict__ gout,      // [B,32,D,H,W]
                       const float *__restrict__ fea,       // [B,V,32,H,W] (MODE_VAR)
                       const float4 *__restrict__ src_cl,    // [B*nsrc][HW][8] (MODE_VAR)
                       const float *__restrict__ rt, const float *__restrict__ depth_values,
                       float *__restrict__ grad_fea,         // [B,V,32,H,W]: view 0 accumulated here (MODE_VAR)
                       float4 *__restrict__ gsrc_cl,         // [B*nsrc][HW][8], zero-initialised
                       int V, int nsrc, int D, int H, int W, int dchunk) {
    __shared__ float4 s_w[kWarps][32];
    __shared__ int4 s_o[kWarps][32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = lane >> 3, g = lane & 7;
    const int nchunks = (D + dchunk - 1) / dchunk;
    const int b = blockIdx.z / nchunks;
    const int d_begin = (blockIdx.z % nchunks) * dchunk;
    const int d_end = min(D, d_begin + dchunk);
    const int y = blockIdx.y * kWarps + warp;
    if (y >= H) return;
    const int x0 = blockIdx.x * 32;
    const int xr = x0 + 8 * p;
    const size_t HW = (size_t)H * W;
    const float invV = 1.0f / (float)V;
    const float twoOverV = 2.0f / (float)V;
    const float xl = (float)(x0 + lane), yf = (float)y;

    float R[8][4], GR[8][4];  // reference-view values and their gradient accumulated over the chunk
    if (MODE == MODE_VAR) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                R[i][j] = (xr + i < W) ? __ldg(fea + (((size_t)b * V) * kC + 4 * g + j) * HW + (size_t)y * W + xr + i) : 0.f;
                GR[i][j] = 0.f;
            }
    }

    for (int d = d_begin; d < d_end; ++d) {
        const float dep = __ldg(depth_values + (size_t)b * D + d);
        float K[8][4];  // MODE_VAR: (2/V) g ; MODE_WARP: g
        float S[8][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float gv = (xr + i < W)
                                     ? __ldg(gout + (((size_t)b * kC + 4 * g + j) * D + d) * HW + (size_t)y * W + xr + i)
                                     : 0.f;
                K[i][j] = (MODE == MODE_VAR) ? twoOverV * gv : gv;
                S[i][j] = (MODE == MODE_VAR) ? R[i][j] : 0.f;
            }

        if (MODE == MODE_VAR) {
            // pass 1: sum over views
            for (int v = 0; v < nsrc; ++v) {
                const int n = b * nsrc + v;
                const Taps t = sample_taps(rt + (size_t)n * 12, xl, yf, dep, H, W);
                __syncwarp();
                s_w[warp][lane] = t.w;
                s_o[warp][lane] = t.off;
                __syncwarp();
                const float4 *f = src_cl + (size_t)n * HW * 8 + g;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 val = bilerp4(f, s_o[warp][8 * p + i], s_w[warp][8 * p + i]);
                    S[i][0] += val.x; S[i][1] += val.y; S[i][2] += val.z; S[i][3] += val.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    S[i][j] *= invV;  // mean
                    GR[i][j] = fmaf(K[i][j], R[i][j] - S[i][j], GR[i][j]);
                }
        }

        // pass 2: per-view gradient, scattered through the taps
        for (int v = 0; v < nsrc; ++v) {
            const int n = b * nsrc + v;
            const Taps t = sample_taps(rt + (size_t)n * 12, xl, yf, dep, H, W);
            __syncwarp();
            s_w[warp][lane] = t.w;
            s_o[warp][lane] = t.off;
            __syncwarp();
            const float4 *f = src_cl + (size_t)n * HW * 8 + g;
            float4 *gf = gsrc_cl + (size_t)n * HW * 8 + g;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (xr + i >= W) continue;
                const float4 w = s_w[warp][8 * p + i];
                const int4 o = s_o[warp][8 * p + i];
                float4 gw;
                if (MODE == MODE_VAR) {
                    const float4 val = bilerp4(f, o, w);
                    gw = make_float4(K[i][0] * (val.x - S[i][0]), K[i][1] * (val.y - S[i][1]),
                                     K[i][2] * (val.z - S[i][2]), K[i][3] * (val.w - S[i][3]));
                } else {
                    gw = make_float4(K[i][0], K[i][1], K[i][2], K[i][3]);
                }
                if (w.x != 0.f) atomicAdd(gf + (size_t)o.x * 8, make_float4(gw.x * w.x, gw.y * w.x, gw.z * w.x, gw.w * w.x));
                if (w.y != 0.f) atomicAdd(gf + (size_t)o.y * 8, make_float4(gw.x * w.y, gw.y * w.y, gw.z * w.y, gw.w * w.y));
                if (w.z != 0.f) atomicAdd(gf + (size_t)o.z * 8, make_float4(gw.x * w.z, gw.y * w.z, gw.z * w.z, gw.w * w.z));
                if (w.w != 0.f) atomicAdd(gf + (size_t)o.w * 8, make_float4(gw.x * w.w, gw.y * w.w, gw.z * w.w, gw.w * w.w));
            }
        }
    }

    if (MODE == MODE_VAR) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (xr + i < W)
                    atomicAdd(grad_fea + (((size_t)b * V) * kC + 4 * g + j) * HW + (size_t)y * W + xr + i, GR[i][j]);
    }
}

// ------------------------------------------------------------------------------------------------
// Host-side launchers
// ------------------------------------------------------------------------------------------------
static inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

int warp_variance_windows_f32(const float *fea, void *tex32, const float *rt, const float *depth_values, float *var, int B, int V,
                              int D, int H, int W, cudaStream_t st);  // warp_variance_win.cu

static int pick_dchunk(int B, int D, int H, int W) {
    // enough CTAs for >= 4 waves of 148 SMs x 2 resident CTAs, but keep depth runs long (L1 reuse)
    const long long tiles = (long long)cdiv(W, 32) * cdiv(H, kWarps) * B;
    int dchunk = 16;
    while (dchunk > 2 && tiles * cdiv(D, dchunk) < 148LL * 2 * 4) dchunk >>= 1;
    while ((long long)B * cdiv(D, dchunk) > 65535) dchunk <<= 1;  // gridDim.z limit
    return dchunk;
}

}  // namespace mvs

using namespace mvs;

extern "C" size_t mvs_warp_variance_workspace_bytes(int B, int V, int C, int H, int W) {
    if (B <= 0 || V < 1 || C <= 0 || H <= 0 || W <= 0) return 0;
    // homographies | channels-last sources (fp32, or bf16 + an fp32 channels-last reference view in the bf16-texel path)
    return align256((size_t)B * (V > 1 ? V - 1 : 1) * 12 * sizeof(float)) +
           align256((size_t)B * (V > 1 ? V - 1 : 1) * H * W * C * sizeof(float)) +
           align256((size_t)B * H * W * C * sizeof(float)) + 512;
}

extern "C" size_t mvs_warp_variance_bwd_workspace_bytes(int B, int V, int C, int H, int W) {
    if (B <= 0 || V < 1 || C <= 0 || H <= 0 || W <= 0) return 0;
    return mvs_warp_variance_workspace_bytes(B, V, C, H, W) +
           align256((size_t)B * (V > 1 ? V - 1 : 1) * H * W * C * sizeof(float));
}

static int check_dims(int B, int C, int D, int H, int W) {
    MVS_REQUIRE(B > 0 && C > 0 && D > 0 && H > 1 && W > 1, "bad shape B=%d C=%d D=%d H=%d W=%d (need H,W >= 2)", B, C, D,
                H, W);
    MVS_REQUIRE((long long)B * D <= 65535LL * 16, "B*D=%lld too large for one launch", (long long)B * D);
    MVS_REQUIRE((long long)H * W < (1LL << 27), "feature map too large");
    return MVS_OK;
}

extern "C" int mvs_warp_variance_fwd(const float *fea, const float *proj, const float *depth_values, float *var,
                                     void *workspace, int B, int V, int C, int D, int H, int W, void *stream) {
    MVS_REQUIRE(fea && proj && depth_values && var && workspace, "null pointer argument");
    MVS_REQUIRE(C == kC, "warp_variance: C must be %d (FeatureNet width), got %d", kC, C);
    MVS_REQUIRE(V >= 1 && V <= 64, "warp_variance: V=%d out of range", V);
    if (int rc = check_dims(B, C, D, H, W)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nsrc = V - 1;
    float *rt = (float *)workspace;
    // behind the homographies: room for every view's feature map in the window kernel's layout (B*V*H*W*128 bytes)
    float *tex32 = (float *)((char *)workspace + align256((size_t)B * (nsrc > 0 ? nsrc : 1) * 12 * sizeof(float)));
    if (nsrc > 0)
        if (int rc = compose_homographies(proj, rt, B, V, st)) return rc;
    // TMA-window kernel (warp_variance_win.cu) for up to 8 source views; beyond that no view gets a window and the older
    // per-tap gather kernel below is the faster of the two (bit-identical results either way)
    if (nsrc <= 8) return warp_variance_windows_f32(fea, tex32, rt, depth_values, var, B, V, D, H, W, st);
    const int HW = H * W;
    nchw_to_nhwc32_kernel<<<dim3(cdiv(HW, 32), B * nsrc), dim3(32, 8), 0, st>>>(fea, tex32, HW, nsrc, V);
    MVS_LAUNCH_CHECK(1);
    const int dchunk = pick_dchunk(B, D, H, W);
    dim3 grid(cdiv(W, 32), cdiv(H, kWarps), B * cdiv(D, dchunk));
    warp_variance_fwd2_kernel<<<grid, kThreads, 0, st>>>(fea, (const float4 *)tex32, rt, depth_values, var, V, nsrc, D, H, W, dchunk);
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

// Internal: same op, output written as fp16 CP8 [B][4][D][H][W][8] for the tensor-core CostRegNet.
namespace mvs {
int warp_variance_windows(const void *tex16, const float *rt, const float *depth_values, void *vol_cp8, int B, int V, int D,
                          int H, int W, cudaStream_t st, int n_images = 0, const int *view_ids_host = nullptr);
int features_nchw_to_rcp8(const float *fea, void *tex16, int N, int H, int W, cudaStream_t st);
int features_nhwc16_to_rcp8(const void *fea16, void *tex16, int N, int H, int W, cudaStream_t st);
// fp32 NCHW features in: one layout pass to fp16 RCP8 texels (all V views), then the TMA-window kernel
// (warp_variance_win.cu).  workspace (sized for fp32 texels): rt | fp16 RCP8 features.
int warp_variance_cp8(const float *fea, const float *proj, const float *depth_values, void *vol_cp8, void *workspace,
                      int B, int V, int D, int H, int W, cudaStream_t st) {
    const int nsrc = V - 1;
    float *rt = (float *)workspace;
    void *tex16 = (char *)workspace + align256((size_t)B * (nsrc > 0 ? nsrc : 1) * 12 * sizeof(float));
    if (nsrc > 0)
        if (int rc = compose_homographies(proj, rt, B, V, st)) return rc;
    if (int rc = features_nchw_to_rcp8(fea, tex16, B * V, H, W, st)) return rc;
    return warp_variance_windows(tex16, rt, depth_values, vol_cp8, B, V, D, H, W, st);
}

// fp16 RCP8 features of all V views in ([B*V][H][4][W][8], what the tensor-core FeatureNet writes): no layout pass.
int warp_variance_cp8_rcp8(const void *tex16, const float *proj, const float *depth_values, void *vol_cp8, void *workspace,
                           int B, int V, int D, int H, int W, cudaStream_t st) {
    float *rt = (float *)workspace;
    if (V > 1)
        if (int rc = compose_homographies(proj, rt, B, V, st)) return rc;
    return warp_variance_windows(tex16, rt, depth_values, vol_cp8, B, V, D, H, W, st);
}

// fp16 RCP8 features of a POOL of images; view v of batch element b is image view_ids_host[b*V + v]
int warp_variance_cp8_pool(const void *pool16, int n_pool, const int *view_ids_host, const float *proj,
                           const float *depth_values, void *vol_cp8, void *workspace, int B, int V, int D, int H, int W,
                           cudaStream_t st) {
    float *rt = (float *)workspace;
    if (V > 1)
        if (int rc = compose_homographies(proj, rt, B, V, st)) return rc;
    return warp_variance_windows(pool16, rt, depth_values, vol_cp8, B, V, D, H, W, st, n_pool, view_ids_host);
}

// fp16 channels-last features of all V views in ([B][V][H*W][32], what a half-precision cuDNN FeatureNet emits)
int warp_variance_cp8_f16(const void *fea16, const float *proj, const float *depth_values, void *vol_cp8, void *workspace,
                          int B, int V, int D, int H, int W, cudaStream_t st) {
    const int nsrc = V - 1;
    float *rt = (float *)workspace;
    if (nsrc > 0)
        if (int rc = compose_homographies(proj, rt, B, V, st)) return rc;
    void *tex16 = (char *)workspace + align256((size_t)B * (nsrc > 0 ? nsrc : 1) * 12 * sizeof(float));
    if (int rc = features_nhwc16_to_rcp8(fea16, tex16, B * V, H, W, st)) return rc;
    return warp_variance_windows(tex16, rt, depth_values, vol_cp8, B, V, D, H, W, st);
}
}  // namespace mvs

extern "C" int mvs_homo_warping(const float *src_fea, const float *src_proj, const float *ref_proj,
                                const float *depth_values, float *out, int B, int C, int D, int H, int W,
                                void *stream) {
    MVS_REQUIRE(src_fea && src_proj && ref_proj && depth_values && out, "null pointer argument");
    if (int rc = check_dims(B, C, D, H, W)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    // scratch: homographies (+ channels-last copy when C == 32); stream-ordered allocation, no host sync
    const size_t rt_bytes = align256((size_t)B * 12 * sizeof(float));
    const size_t cl_bytes = (C == kC) ? (size_t)B * HW * kC * sizeof(float) : 0;
    void *ws = nullptr;
    MVS_CUDA(cudaMallocAsync(&ws, rt_bytes + cl_bytes, st));
    float *rt = (float *)ws;
    int rc = compose_homography_pairs(src_proj, ref_proj, rt, B, st);
    if (rc == MVS_OK) {
        if (C == kC) {
            float *src_cl = (float *)((char *)ws + rt_bytes);
            nchw_to_nhwc32_kernel<<<dim3(cdiv(HW, 32), B), dim3(32, 8), 0, st>>>(src_fea, src_cl, HW, 1, 1);
            const int dchunk = pick_dchunk(B, D, H, W);
            dim3 grid(cdiv(W, 32), cdiv(H, kWarps), B * cdiv(D, dchunk));
            warp_volume_fwd_kernel<MODE_WARP><<<grid, kThreads, 0, st>>>(nullptr, (const float4 *)src_cl, rt,
                                                                         depth_values, out, 1, 1, D, H, W, dchunk);
            count_launches(2);
        } else {
            dim3 grid(cdiv(W, 128), H, B * D);
            homo_warp_generic_kernel<<<grid, 128, 0, st>>>(src_fea, rt, depth_values, out, C, D, H, W);
            count_launches(1);
        }
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(ws, st);
    if (rc != MVS_OK) return rc;
    if (e != cudaSuccess) return set_error(MVS_ERR_CUDA, "homo_warping launch failed: %s", cudaGetErrorString(e));
    return MVS_OK;
}

extern "C" int mvs_homo_warping_bwd(const float *grad_out, const float *src_proj, const float *ref_proj,
                                    const float *depth_values, float *grad_src, int B, int C, int D, int H, int W,
                                    void *stream) {
    MVS_REQUIRE(grad_out && src_proj && ref_proj && depth_values && grad_src, "null pointer argument");
    if (int rc = check_dims(B, C, D, H, W)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    const size_t rt_bytes = align256((size_t)B * 12 * sizeof(float));
    const size_t cl_bytes = (C == kC) ? (size_t)B * HW * kC * sizeof(float) : 0;
    void *ws = nullptr;
    MVS_CUDA(cudaMallocAsync(&ws, rt_bytes + cl_bytes, st));
    float *rt = (float *)ws;
    int rc = compose_homography_pairs(src_proj, ref_proj, rt, B, st);
    cudaError_t e = cudaSuccess;
    if (rc == MVS_OK) {
        if (C == kC) {
            float *g_cl = (float *)((char *)ws + rt_bytes);
            e = cudaMemsetAsync(g_cl, 0, cl_bytes, st);
            const int dchunk = pick_dchunk(B, D, H, W);
            dim3 grid(cdiv(W, 32), cdiv(H, kWarps), B * cdiv(D, dchunk));
            warp_volume_bwd_kernel<MODE_WARP><<<grid, kThreads, 0, st>>>(grad_out, nullptr, nullptr, rt, depth_values,
                                                                         nullptr, (float4 *)g_cl, 1, 1, D, H, W, dchunk);
            nhwc32_to_nchw_kernel<<<dim3(cdiv(HW, 32), B), dim3(32, 8), 0, st>>>(g_cl, grad_src, HW, 1, 1);
            count_launches(2);
        } else {
            e = cudaMemsetAsync(grad_src, 0, (size_t)B * C * HW * sizeof(float), st);
            dim3 grid(cdiv(W, 128), H, B * D);
            homo_warp_generic_bwd_kernel<<<grid, 128, 0, st>>>(grad_out, rt, depth_values, grad_src, C, D, H, W);
            count_launches(1);
        }
    }
    cudaError_t e2 = cudaGetLastError();
    cudaFreeAsync(ws, st);
    if (rc != MVS_OK) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess)
        return set_error(MVS_ERR_CUDA, "homo_warping_bwd failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return MVS_OK;
}

extern "C" int mvs_warp_variance_bwd(const float *grad_var, const float *fea, const float *proj,
                                     const float *depth_values, float *grad_fea, void *workspace, int B, int V, int C,
                                     int D, int H, int W, void *stream) {
    MVS_REQUIRE(grad_var && fea && proj && depth_values && grad_fea && workspace, "null pointer argument");
    MVS_REQUIRE(C == kC, "warp_variance_bwd: C must be %d, got %d", kC, C);
    MVS_REQUIRE(V >= 1 && V <= 64, "warp_variance_bwd: V=%d out of range", V);
    if (int rc = check_dims(B, C, D, H, W)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nsrc = V - 1;
    const int HW = H * W;
    const size_t rt_bytes = align256((size_t)B * (nsrc > 0 ? nsrc : 1) * 12 * sizeof(float));
    const size_t cl_bytes = align256((size_t)B * (nsrc > 0 ? nsrc : 1) * HW * kC * sizeof(float));
    float *rt = (float *)workspace;
    float *src_cl = (float *)((char *)workspace + rt_bytes);
    float *g_cl = (float *)((char *)workspace + rt_bytes + cl_bytes);
    // grad of the reference view is accumulated with atomics across depth chunks: start from zero
    for (int b = 0; b < B; ++b)
        MVS_CUDA(cudaMemsetAsync(grad_fea + (size_t)b * V * kC * HW, 0, (size_t)kC * HW * sizeof(float), st));
    if (nsrc > 0) {
        MVS_CUDA(cudaMemsetAsync(g_cl, 0, (size_t)B * nsrc * HW * kC * sizeof(float), st));
        if (int rc = compose_homographies(proj, rt, B, V, st)) return rc;
        nchw_to_nhwc32_kernel<<<dim3(cdiv(HW, 32), B * nsrc), dim3(32, 8), 0, st>>>(fea, src_cl, HW, nsrc, V);
        MVS_LAUNCH_CHECK(1);
    }
    const int dchunk = pick_dchunk(B, D, H, W);
    dim3 grid(cdiv(W, 32), cdiv(H, kWarps), B * cdiv(D, dchunk));
    warp_volume_bwd_kernel<MODE_VAR><<<grid, kThreads, 0, st>>>(grad_var, fea, (const float4 *)src_cl, rt, depth_values,
                                                                grad_fea, (float4 *)g_cl, V, nsrc, D, H, W, dchunk);
    MVS_LAUNCH_CHECK(1);
    if (nsrc > 0) {
        nhwc32_to_nchw_kernel<<<dim3(cdiv(HW, 32), B * nsrc), dim3(32, 8), 0, st>>>(g_cl, grad_fea, HW, nsrc, V);
        MVS_LAUNCH_CHECK(1);
    }
    return MVS_OK;
}
