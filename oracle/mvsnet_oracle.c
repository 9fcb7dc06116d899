/*
 * mvsnet_oracle.c -- CPU restatement of the MVSNet depth-inference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or the timed CPU baseline.
 *
 * Every function restates, in plain C, the arithmetic the reference performs
 * through PyTorch library calls; the reference file:line each one follows is
 * cited on the function.  The arithmetic itself lives in a third-party
 * dependency of the reference (PyTorch: requirements.txt:1 "torch", unpinned;
 * README.md:26 says 2.0.1; 2.11.0+cu128 is what is installed here), so the
 * sampler semantics are restated from its published behaviour:
 *   - F.grid_sample(bilinear, zeros, align_corners=False)
 *       un-normalise  ix = ((gx + 1) * W - 1) / 2          (ATen GridSampler.h)
 *       taps floor(ix)+{0,1}, floor(iy)+{0,1}; a tap contributes only when it
 *       lies inside [0,W) x [0,H); non-finite / out-of-int-range coordinates
 *       map to -100 (the CUDA kernel's rule, GridSampler.cuh) => no sample.
 *   - nn.Conv3d / nn.ConvTranspose3d / nn.BatchNorm3d(eval) / F.softmax.
 *
 * Pinning: the reference ships no tests or golden vectors, so
 * this oracle is pinned against outputs of the unmodified reference imported
 * in the build container (tests/golden/ npz files, made by tests/make_golden.py).
 *
 * Layouts are the reference's: features NCHW, volumes NCDHW, fp32 everywhere.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* models/module.py:107-109   proj = src_proj @ inverse(ref_proj); rot, trans  */
/* The reference does this in fp32 through LAPACK; here the inverse and the    */
/* product are taken in double and rounded once, which is what the CUDA side   */
/* does too (csrc/warp_variance.cu: compose_homography_kernel).                */
/* ------------------------------------------------------------------------- */
static int inv4x4(const double m[16], double inv[16]) {
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
        if (a[p][c] == 0.0) return -1;
        if (p != c)
            for (int j = 0; j < 8; ++j) { double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
        double d = 1.0 / a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            double f = a[r][c];
            if (f != 0.0)
                for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) inv[i * 4 + j] = a[i][j + 4];
    return 0;
}

ORC_API int orc_compose_homography(const float *src_proj, const float *ref_proj, float *rot, float *trans) {
    double s[16], r[16], ri[16];
    for (int i = 0; i < 16; ++i) { s[i] = src_proj[i]; r[i] = ref_proj[i]; }
    if (inv4x4(r, ri)) return -1;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += s[i * 4 + k] * ri[k * 4 + j];
            if (j < 3) rot[i * 3 + j] = (float)acc;
            else trans[i] = (float)acc;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* models/module.py:119-136   plane-sweep sample position for one (x, y, d).   */
/* Same fp32 operation order as the reference: R*(x,y,1) -> *depth -> +t ->    */
/* /z -> /((W-1)/2) - 1 -> grid_sample un-normalise with align_corners=False.  */
/* volatile-free, but compiled with -ffp-contract=off so no FMA contraction.   */
/* ------------------------------------------------------------------------- */
static inline float safe_coord(float v) {
    /* GridSampler.cuh safe_downgrade_to_int_range */
    if (!(v <= 2147483520.0f) || !(v >= -2147483648.0f) || !isfinite(v)) return -100.0f;
    return v;
}

static inline void sample_pos(const float *rot, const float *trans, float x, float y, float d, int H, int W,
                              float *ix, float *iy) {
    float rx = rot[0] * x + rot[1] * y + rot[2] * 1.0f; /* matmul(rot, xyz)  :125 */
    float ry = rot[3] * x + rot[4] * y + rot[5] * 1.0f;
    float rz = rot[6] * x + rot[7] * y + rot[8] * 1.0f;
    float qx = rx * d + trans[0]; /* :126-128 */
    float qy = ry * d + trans[1];
    float qz = rz * d + trans[2];
    float px = qx / qz; /* :129 */
    float py = qy / qz;
    float gx = px / ((float)(W - 1) / 2.0f) - 1.0f; /* :130 */
    float gy = py / ((float)(H - 1) / 2.0f) - 1.0f; /* :131 */
    *ix = safe_coord(((gx + 1.0f) * (float)W - 1.0f) / 2.0f); /* grid_sample, align_corners=False */
    *iy = safe_coord(((gy + 1.0f) * (float)H - 1.0f) / 2.0f);
}

/* models/module.py:96-139  homo_warping: [B,C,H,W] -> [B,C,D,H,W] */
ORC_API void orc_homo_warp(const float *src_fea, const float *rot, const float *trans, const float *depth_values,
                           float *out, int B, int C, int D, int H, int W) {
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d) {
            const float *R = rot + b * 9, *T = trans + b * 3;
            const float dep = depth_values[b * D + d];
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float ix, iy;
                    sample_pos(R, T, (float)x, (float)y, dep, H, W, &ix, &iy);
                    float fx0 = floorf(ix), fy0 = floorf(iy);
                    int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
                    float wnw = ((fx0 + 1.0f) - ix) * ((fy0 + 1.0f) - iy);
                    float wne = (ix - fx0) * ((fy0 + 1.0f) - iy);
                    float wsw = ((fx0 + 1.0f) - ix) * (iy - fy0);
                    float wse = (ix - fx0) * (iy - fy0);
                    int vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W;
                    int vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
                    for (int c = 0; c < C; ++c) {
                        const float *f = src_fea + ((size_t)b * C + c) * HW;
                        float v = 0.0f;
                        if (vx0 && vy0) v += f[(size_t)y0 * W + x0] * wnw;
                        if (vx1 && vy0) v += f[(size_t)y0 * W + x1] * wne;
                        if (vx0 && vy1) v += f[(size_t)y1 * W + x0] * wsw;
                        if (vx1 && vy1) v += f[(size_t)y1 * W + x1] * wse;
                        out[(((size_t)b * C + c) * D + d) * HW + (size_t)y * W + x] = v;
                    }
                }
        }
}

/* ------------------------------------------------------------------------- */
/* models/mvsnet.py:145-177  variance cost volume over V views.                */
/* fea [B,V,C,H,W] (view 0 = reference view), rot [B,V-1,9], trans [B,V-1,3].  */
/* S = ref; Q = ref^2; per source view S += w, Q += w^2; var = Q/V - (S/V)^2.  */
/* ------------------------------------------------------------------------- */
ORC_API void orc_warp_variance(const float *fea, const float *rot, const float *trans, const float *depth_values,
                               float *var, int B, int V, int C, int D, int H, int W) {
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d) {
            float *S = (float *)malloc(sizeof(float) * C * 2);
            float *Q = S + C;
            const float dep = depth_values[b * D + d];
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const size_t pix = (size_t)y * W + x;
                    for (int c = 0; c < C; ++c) {
                        float r = fea[(((size_t)b * V + 0) * C + c) * HW + pix];
                        S[c] = r;
                        Q[c] = r * r;
                    }
                    for (int v = 1; v < V; ++v) {
                        const float *R = rot + ((size_t)b * (V - 1) + (v - 1)) * 9;
                        const float *T = trans + ((size_t)b * (V - 1) + (v - 1)) * 3;
                        float ix, iy;
                        sample_pos(R, T, (float)x, (float)y, dep, H, W, &ix, &iy);
                        float fx0 = floorf(ix), fy0 = floorf(iy);
                        int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
                        float wnw = ((fx0 + 1.0f) - ix) * ((fy0 + 1.0f) - iy);
                        float wne = (ix - fx0) * ((fy0 + 1.0f) - iy);
                        float wsw = ((fx0 + 1.0f) - ix) * (iy - fy0);
                        float wse = (ix - fx0) * (iy - fy0);
                        int vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W;
                        int vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
                        for (int c = 0; c < C; ++c) {
                            const float *f = fea + (((size_t)b * V + v) * C + c) * HW;
                            float w = 0.0f;
                            if (vx0 && vy0) w += f[(size_t)y0 * W + x0] * wnw;
                            if (vx1 && vy0) w += f[(size_t)y0 * W + x1] * wne;
                            if (vx0 && vy1) w += f[(size_t)y1 * W + x0] * wsw;
                            if (vx1 && vy1) w += f[(size_t)y1 * W + x1] * wse;
                            S[c] = S[c] + w;
                            Q[c] = Q[c] + w * w;
                        }
                    }
                    for (int c = 0; c < C; ++c) {
                        float m = S[c] / (float)V;
                        var[(((size_t)b * C + c) * D + d) * HW + pix] = Q[c] / (float)V - m * m;
                    }
                }
            free(S);
        }
}

/* ------------------------------------------------------------------------- */
/* Backward of the block above (autograd through mvsnet.py:167-177 and         */
/* grid_sample's input gradient; the grid is built under no_grad,              */
/* module.py:106, so only the features receive gradient).                      */
/*   dL/dx_v = (2/V) (x_v - mean) g   for every view, reference included;      */
/* reference-view grad sums over D, source grads scatter through the 4 taps.   */
/* grad_fea [B,V,C,H,W] is fully written (zero-initialised here).              */
/* Accumulation is in double so the oracle is order-independent.               */
/* ------------------------------------------------------------------------- */
ORC_API void orc_warp_variance_bwd(const float *grad_var, const float *fea, const float *rot, const float *trans,
                                   const float *depth_values, float *grad_fea, int B, int V, int C, int D, int H,
                                   int W) {
    const size_t HW = (size_t)H * W;
    const size_t n = (size_t)B * V * C * HW;
    double *acc = (double *)calloc(n, sizeof(double));
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            float *wv = (float *)malloc(sizeof(float) * V);
            for (int d = 0; d < D; ++d) {
                const float dep = depth_values[b * D + d];
                for (int y = 0; y < H; ++y)
                    for (int x = 0; x < W; ++x) {
                        const size_t pix = (size_t)y * W + x;
                        float S = fea[(((size_t)b * V + 0) * C + c) * HW + pix];
                        wv[0] = S;
                        int x0s[16], y0s[16];
                        float wts[16][4];
                        for (int v = 1; v < V; ++v) {
                            const float *R = rot + ((size_t)b * (V - 1) + (v - 1)) * 9;
                            const float *T = trans + ((size_t)b * (V - 1) + (v - 1)) * 3;
                            float ix, iy;
                            sample_pos(R, T, (float)x, (float)y, dep, H, W, &ix, &iy);
                            float fx0 = floorf(ix), fy0 = floorf(iy);
                            int x0 = (int)fx0, y0 = (int)fy0;
                            x0s[v] = x0; y0s[v] = y0;
                            wts[v][0] = ((fx0 + 1.0f) - ix) * ((fy0 + 1.0f) - iy);
                            wts[v][1] = (ix - fx0) * ((fy0 + 1.0f) - iy);
                            wts[v][2] = ((fx0 + 1.0f) - ix) * (iy - fy0);
                            wts[v][3] = (ix - fx0) * (iy - fy0);
                            const float *f = fea + (((size_t)b * V + v) * C + c) * HW;
                            float w = 0.0f;
                            for (int t = 0; t < 4; ++t) {
                                int xx = x0 + (t & 1), yy = y0 + (t >> 1);
                                if (xx >= 0 && xx < W && yy >= 0 && yy < H) w += f[(size_t)yy * W + xx] * wts[v][t];
                            }
                            wv[v] = w;
                            S += w;
                        }
                        float mean = S / (float)V;
                        float g = grad_var[(((size_t)b * C + c) * D + d) * HW + pix];
                        float k = 2.0f / (float)V * g;
                        acc[(((size_t)b * V + 0) * C + c) * HW + pix] += (double)(k * (wv[0] - mean));
                        for (int v = 1; v < V; ++v) {
                            float gw = k * (wv[v] - mean);
                            double *ga = acc + (((size_t)b * V + v) * C + c) * HW;
                            for (int t = 0; t < 4; ++t) {
                                int xx = x0s[v] + (t & 1), yy = y0s[v] + (t >> 1);
                                if (xx >= 0 && xx < W && yy >= 0 && yy < H)
                                    ga[(size_t)yy * W + xx] += (double)(gw * wts[v][t]);
                            }
                        }
                    }
            }
            free(wv);
        }
    for (size_t i = 0; i < n; ++i) grad_fea[i] = (float)acc[i];
    free(acc);
}

/* ------------------------------------------------------------------------- */
/* models/module.py:26-33 (ConvBnReLU3D) and mvsnet.py:62 (prob conv).         */
/* nn.Conv3d k=3 pad=1, stride 1 or 2, weight [Cout,Cin,3,3,3], optional bias; */
/* optional eval-mode BatchNorm3d  y = (x-mean)/sqrt(var+eps)*gamma+beta       */
/* (bn = {gamma,beta,mean,var} each [Cout], NULL to skip); optional ReLU.      */
/* Accumulation in double: the oracle is the "true" value for both the fp32    */
/* and the bf16 product kernels.                                               */
/* ------------------------------------------------------------------------- */
ORC_API void orc_conv3d(const float *in, const float *wgt, const float *bias, const float *bn_gamma,
                        const float *bn_beta, const float *bn_mean, const float *bn_var, float eps, int relu,
                        float *out, int B, int Cin, int Cout, int Din, int Hin, int Win, int stride) {
    const int Do = (Din + 2 - 3) / stride + 1, Ho = (Hin + 2 - 3) / stride + 1, Wo = (Win + 2 - 3) / stride + 1;
    const size_t in_cs = (size_t)Din * Hin * Win, out_cs = (size_t)Do * Ho * Wo;
#pragma omp parallel for collapse(3) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int co = 0; co < Cout; ++co)
            for (int z = 0; z < Do; ++z) {
                double *row = (double *)malloc(sizeof(double) * Wo);
                for (int y = 0; y < Ho; ++y) {
                    for (int x = 0; x < Wo; ++x) row[x] = bias ? (double)bias[co] : 0.0;
                    for (int ci = 0; ci < Cin; ++ci) {
                        const float *ip = in + ((size_t)b * Cin + ci) * in_cs;
                        const float *wp = wgt + ((size_t)co * Cin + ci) * 27;
                        for (int kd = 0; kd < 3; ++kd) {
                            int iz = z * stride - 1 + kd;
                            if (iz < 0 || iz >= Din) continue;
                            for (int kh = 0; kh < 3; ++kh) {
                                int iy = y * stride - 1 + kh;
                                if (iy < 0 || iy >= Hin) continue;
                                const float *irow = ip + ((size_t)iz * Hin + iy) * Win;
                                for (int kw = 0; kw < 3; ++kw) {
                                    double w = wp[(kd * 3 + kh) * 3 + kw];
                                    for (int x = 0; x < Wo; ++x) {
                                        int ixx = x * stride - 1 + kw;
                                        if (ixx >= 0 && ixx < Win) row[x] += w * (double)irow[ixx];
                                    }
                                }
                            }
                        }
                    }
                    float *op = out + (((size_t)b * Cout + co) * Do + z) * (size_t)Ho * Wo + (size_t)y * Wo;
                    for (int x = 0; x < Wo; ++x) {
                        float v = (float)row[x];
                        if (bn_gamma) v = (v - bn_mean[co]) / sqrtf(bn_var[co] + eps) * bn_gamma[co] + bn_beta[co];
                        if (relu && v < 0.0f) v = 0.0f;
                        op[x] = v;
                    }
                }
                free(row);
            }
    (void)out_cs;
}

/* ------------------------------------------------------------------------- */
/* mvsnet.py:46-59  nn.ConvTranspose3d(k=3, stride=2, padding=1,               */
/* output_padding=1, bias=False) + BatchNorm3d(eval) + ReLU, then the skip add */
/* of mvsnet.py:69-71 (skip may be NULL).  weight [Cin,Cout,3,3,3].            */
/* out[o] = sum over i,k with o = 2 i - 1 + k.  Output dims are 2x the input.  */
/* ------------------------------------------------------------------------- */
ORC_API void orc_convT3d(const float *in, const float *wgt, const float *bn_gamma, const float *bn_beta,
                         const float *bn_mean, const float *bn_var, float eps, int relu, const float *skip,
                         float *out, int B, int Cin, int Cout, int Din, int Hin, int Win) {
    const int Do = 2 * Din, Ho = 2 * Hin, Wo = 2 * Win;
    const size_t in_cs = (size_t)Din * Hin * Win;
#pragma omp parallel for collapse(3) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int co = 0; co < Cout; ++co)
            for (int z = 0; z < Do; ++z)
                for (int y = 0; y < Ho; ++y)
                    for (int x = 0; x < Wo; ++x) {
                        double acc = 0.0;
                        for (int kd = 0; kd < 3; ++kd) {
                            int tz = z + 1 - kd;
                            if (tz < 0 || (tz & 1) || tz / 2 >= Din) continue;
                            for (int kh = 0; kh < 3; ++kh) {
                                int ty = y + 1 - kh;
                                if (ty < 0 || (ty & 1) || ty / 2 >= Hin) continue;
                                for (int kw = 0; kw < 3; ++kw) {
                                    int tx = x + 1 - kw;
                                    if (tx < 0 || (tx & 1) || tx / 2 >= Win) continue;
                                    size_t ioff = ((size_t)(tz / 2) * Hin + ty / 2) * Win + tx / 2;
                                    for (int ci = 0; ci < Cin; ++ci)
                                        acc += (double)in[((size_t)b * Cin + ci) * in_cs + ioff] *
                                               (double)wgt[(((size_t)ci * Cout + co) * 27) + (kd * 3 + kh) * 3 + kw];
                                }
                            }
                        }
                        float v = (float)acc;
                        if (bn_gamma) v = (v - bn_mean[co]) / sqrtf(bn_var[co] + eps) * bn_gamma[co] + bn_beta[co];
                        if (relu && v < 0.0f) v = 0.0f;
                        size_t o = ((((size_t)b * Cout + co) * Do + z) * Ho + y) * (size_t)Wo + x;
                        if (skip) v = skip[o] + v;
                        out[o] = v;
                    }
}

/* ------------------------------------------------------------------------- */
/* mvsnet.py:192-193 softmax over depth; module.py:144-147 depth_regression;   */
/* mvsnet.py:214-218 photometric confidence:                                   */
/*   sum4[d] = p[d-1] + p[d] + p[d+1] + p[d+2]   (zero padded 1 front, 2 back) */
/*   i = trunc(sum_d p[d] * d)   (.long());  conf = sum4[i]                    */
/* logits [B,D,H,W]; depth_values [B,D]; prob may be NULL.                     */
/* ------------------------------------------------------------------------- */
ORC_API void orc_softmax_depth_conf(const float *logits, const float *depth_values, float *prob, float *depth,
                                    float *conf, float *index_f, int B, int D, int H, int W) {
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y) {
            float *p = (float *)malloc(sizeof(float) * D);
            for (int x = 0; x < W; ++x) {
                const float *l = logits + (size_t)b * D * HW + (size_t)y * W + x;
                float m = -INFINITY;
                for (int d = 0; d < D; ++d) m = fmaxf(m, l[d * HW]);
                float s = 0.0f;
                for (int d = 0; d < D; ++d) { p[d] = expf(l[d * HW] - m); s += p[d]; }
                float dep = 0.0f, idx = 0.0f;
                for (int d = 0; d < D; ++d) {
                    p[d] = p[d] / s;
                    dep += p[d] * depth_values[b * D + d];
                    idx += p[d] * (float)d;
                }
                long i = (long)idx;
                if (i < 0) i = 0;
                if (i > D - 1) i = D - 1;
                float c4 = 0.0f;
                for (long k = i - 1; k <= i + 2; ++k)
                    if (k >= 0 && k < D) c4 += p[k];
                size_t o = (size_t)b * HW + (size_t)y * W + x;
                depth[o] = dep;
                conf[o] = c4;
                if (index_f) index_f[o] = idx;
                if (prob)
                    for (int d = 0; d < D; ++d) prob[(size_t)b * D * HW + d * HW + (size_t)y * W + x] = p[d];
            }
            free(p);
        }
}

/* module.py:144-147  depth_regression(p, depth_values); dv_stride = D for [B,D], 0 for a shared [D] vector */
ORC_API void orc_depth_regression(const float *p, const float *depth_values, int dv_stride, float *out, int B,
                                  int D, int H, int W) {
    const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (size_t i = 0; i < HW; ++i) {
            float acc = 0.0f;
            for (int d = 0; d < D; ++d) acc += p[((size_t)b * D + d) * HW + i] * depth_values[b * dv_stride + d];
            out[(size_t)b * HW + i] = acc;
        }
}
