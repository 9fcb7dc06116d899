// Fused plane-sweep warp + variance, third generation: TMA-staged source WINDOWS in shared memory.
//
// Replaces, for the reference (olivier-2018/scene_3Dreconstruction_MVSNet), in the tensor-core precision modes:
//   models/module.py:96-139   homo_warping  (grid construction + F.grid_sample, bilinear, zero padding)
//   models/mvsnet.py:145-177  running sum / sum of squares over views and the variance
//
// Why (ncu on the second generation, profiles/r01e): ~120 instructions per (pixel, view, 8 channels) of which only 32
// are arithmetic -- the rest is tap clamping and masking, 64-bit address arithmetic for 16 global loads, the smem
// exchange of coordinates between the lane that computes them and the lanes that gather, and L2-latency stalls at 16
// warps per SM.  This version removes those instead of hiding them:
//   * Features of ALL views are stored as fp16 "RCP8" [n][y][chunk 0..3][x][8 ch]: a row of one 8-channel chunk is
//     contiguous, so a TMA box over (x, chunk, y) lands in shared memory as [row][chunk][col][16 B].
//   * A CTA owns a TW x TH pixel tile and a chunk of depth planes.  For a run ("segment") of planes, one elected warp
//     bounds the source footprint of the tile in every source view from the 8 corners of (tile x depth range) -- the
//     map (x*d, y*d, d) -> (u, v) is projective, so the corners' bounding box contains every sample while q_z > 0 --
//     and one thread issues one TMA box per view at that origin.  TMA zero-fills outside the image, which IS
//     grid_sample's padding_mode='zeros': no clamping, no validity masks.
//   * Thread = pixel, all 32 channels: coordinates are computed by the thread that uses them (no exchange), the 16
//     16-byte tap loads of a (pixel, view, plane) are LDS.128 at compile-time offsets from ONE 32-bit address, lanes of
//     a warp read consecutive 16-byte columns (conflict-free), Sum / Sum^2 of 32 channels live in 64 registers and the
//     bf16 CP8 output row is written with fully coalesced 16-byte stores.
//   * Generality: if the footprint of a segment does not fit the window, the segment is halved; a single plane that
//     still does not fit (extreme zoom, q_z <= 0 inside the tile, depth <= 0) takes a per-tap global gather with
//     explicit clamps for that view -- slow, exact, never taken on camera-like geometry.  Every sample also checks that
//     its 2x2 footprint is inside the window (memory safety for non-finite coordinates).
// HBM traffic stays the algorithmic minimum (features once, volume once); the window fill is L2 -> smem traffic.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace mvs {

namespace {

constexpr int kC = 32;

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ __half2 as_half2(uint32_t u) { return *reinterpret_cast<const __half2 *>(&u); }

// ix = px * W/(W-1) - 0.5 is module.py:130-131 composed with grid_sample's align_corners=False un-normalisation
// (closed form of the reference's chain; differs from it by ~1e-7 relative, far below the fp16 texel quantisation).
struct Coord {
    float ix, iy;
};
__device__ __forceinline__ Coord project(const float4 c0, const float4 c1, const float4 c2, float xf, float yf, float dep,
                                         float sx, float sy) {
    const float rx = fmaf(c0.x, xf, fmaf(c0.y, yf, c0.z));
    const float ry = fmaf(c1.x, xf, fmaf(c1.y, yf, c1.z));
    const float rz = fmaf(c2.x, xf, fmaf(c2.y, yf, c2.z));
    const float qx = fmaf(rx, dep, c0.w), qy = fmaf(ry, dep, c1.w), qz = fmaf(rz, dep, c2.w);
    const float iz = rcp_approx(qz);
    Coord c;
    c.ix = fmaf(qx * iz, sx, -0.5f);
    c.iy = fmaf(qy * iz, sy, -0.5f);
    return c;
}

// one 8-channel chunk of one (pixel, view): 4 taps -> packed-half interpolation -> fp32 Sum / Sum^2
__device__ __forceinline__ void accumulate_chunk(const uint4 a, const uint4 b, const uint4 c, const uint4 d, const __half2 h00,
                                                 const __half2 h01, const __half2 h10, const __half2 h11, float2 *S,
                                                 float2 *Q) {
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
    const uint32_t wc[4] = {c.x, c.y, c.z, c.w}, wd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2 vh =
            __hfma2(as_half2(wd[j]), h11, __hfma2(as_half2(wc[j]), h10, __hfma2(as_half2(wb[j]), h01, __hmul2(as_half2(wa[j]), h00))));
        // mixed-precision accumulate (FHADD / FHFMA: fp32 += f16, fp32 += f16 * f16, the half taken from either half of the
        // register): no conversion instructions, same results as convert + fp32 add / fma (f16 -> f32 is exact)
        const unsigned short lo = __half_as_ushort(__low2half(vh)), hi = __half_as_ushort(__high2half(vh));
        asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(S[j].x) : "h"(lo));
        asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(S[j].y) : "h"(hi));
        asm("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(Q[j].x) : "h"(lo));
        asm("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(Q[j].y) : "h"(hi));
    }
}

// Deviation form (HACC): the variance is shift invariant, so the running sums are kept over d = x_v - x_ref (zero for
// the reference view itself) instead of x_v.  The subtraction is free -- the interpolation chain starts from -x_ref --
// and the deviations are small exactly where the cost volume matters (matching pixels), so Sum d and Sum d^2 can stay
// in packed half: 24 heavy-pipe instructions per 8 channels instead of 32 and one conversion per plane instead of one
// per view.  Relative error of the sums 2^-11, the same class as the fp16 interpolation itself; |d| must stay below
// 255 / sqrt(V-1) for d^2 not to overflow fp16 (features are O(1)).
__device__ __forceinline__ void accumulate_chunk_dev(const uint4 a, const uint4 b, const uint4 c, const uint4 d, const __half2 h00,
                                                     const __half2 h01, const __half2 h10, const __half2 h11,
                                                     const uint4 nref, __half2 *S, __half2 *Q) {
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
    const uint32_t wc[4] = {c.x, c.y, c.z, c.w}, wd[4] = {d.x, d.y, d.z, d.w};
    const uint32_t wr[4] = {nref.x, nref.y, nref.z, nref.w};  // -x_ref
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2 dv = __hfma2(as_half2(wd[j]), h11,
                                   __hfma2(as_half2(wc[j]), h10, __hfma2(as_half2(wb[j]), h01, __hfma2(as_half2(wa[j]), h00, as_half2(wr[j])))));
        S[j] = __hadd2(S[j], dv);
        Q[j] = __hfma2(dv, dv, Q[j]);
    }
}

// fp32 NCHW [B,V,32,HW] -> fp16 RCP8 [B*V][H][4][W][8]; one thread per output 16-byte chunk
__global__ void nchw_to_rcp8_kernel(const float *__restrict__ in, uint4 *__restrict__ out, int H, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W);
    long long r = i / W;
    const int c = (int)(r & 3);
    r >>= 2;
    const int y = (int)(r % H);
    const long long n = r / H;
    const size_t HW = (size_t)H * W;
    const float *src = in + ((size_t)n * kC + 8 * c) * HW + (size_t)y * W + x;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float lo = fminf(fmaxf(__ldg(src + (size_t)(2 * j) * HW), -65504.f), 65504.f);
        const float hi = fminf(fmaxf(__ldg(src + (size_t)(2 * j + 1) * HW), -65504.f), 65504.f);
        const __half2 h = __floats2half2_rn(lo, hi);
        w[j] = *reinterpret_cast<const uint32_t *>(&h);
    }
    out[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// fp16 NHWC [N][H][W][32] -> fp16 RCP8 [N][H][4][W][8]
__global__ void nhwc16_to_rcp8_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W);
    const long long r = i / W;  // (n*H + y)*4 + c
    const int c = (int)(r & 3);
    out[i] = __ldg(in + ((r >> 2) * W + x) * 4 + c);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// grid = (ceil(W / TW), ceil(H / TH), B * ceil(D / dchunk)), block = 32 * TWW * TH threads (warp = 32 consecutive x of a row)
// dynamic smem: nwin windows of WY rows x [4 chunks][WX cols] x 16 B | homographies | window origins | depths | mbarrier
// ------------------------------------------------------------------------------------------------
template <int TWW, int TH, int WX, bool HACC>
__global__ void __launch_bounds__(32 * TWW * TH, (TWW * TH <= 8) ? 2 : 1)
warp_variance_win_kernel(const __grid_constant__ CUtensorMap tmap,      // fp16 RCP8 features, all views
                         const uint4 *__restrict__ tex,                  // the same memory
                         const float *__restrict__ rt,                   // [B*nsrc][12] rot(9) | trans(3)
                         const float *__restrict__ depth_values,         // [B,D]
                         uint4 *__restrict__ out,                        // bf16 CP8 [B,4,D,H,W,8]
                         int V, int nsrc, int nwin, int D, int H, int W, int dchunk, int WY) {
    // the next kernel (conv0, launched with programmatic stream serialization) may start its set-up on SMs this grid has left
    ptx::pdl_launch_dependents();
    ptx::pdl_wait();  // launched with programmatic stream serialization: homographies and features are read below
    constexpr int TW = 32 * TWW;
    constexpr int ROWQ = 4 * WX;     // uint4 per window row
    constexpr int ROWB = ROWQ * 16;  // bytes per window row
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t win_bytes = (uint32_t)WY * ROWB;
    unsigned char *sp = smem_raw + (size_t)nwin * win_bytes;
    float4 *s_rt = reinterpret_cast<float4 *>(sp);  // [nsrc][3]: (r0 r1 r2 t) per output coordinate
    sp += (size_t)nsrc * 48;
    int2 *s_org = reinterpret_cast<int2 *>(sp);  // [nwin] window origin (texel column / row of window element 0)
    sp += (size_t)nwin * 8;
    int *s_mode = reinterpret_cast<int *>(sp);  // [nwin] 0 = window, 1 = global gather (this segment)
    sp += (size_t)nwin * 4;
    float *s_dep = reinterpret_cast<float *>(sp);  // [dchunk]
    sp += (size_t)dchunk * 4;
    sp = reinterpret_cast<unsigned char *>(((uintptr_t)sp + 15) & ~(uintptr_t)15);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(sp);
    int *s_seg = reinterpret_cast<int *>(sp + 8);  // [0] = planes in the current segment

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (D + dchunk - 1) / dchunk;
    const int b = blockIdx.z / nchunks;
    const int d_begin = (blockIdx.z % nchunks) * dchunk;
    const int d_end = min(D, d_begin + dchunk);
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int x = tx0 + (warp % TWW) * 32 + lane, y = ty0 + warp / TWW;
    const bool live = (x < W) & (y < H);
    const float xf = (float)x, yf = (float)y;
    const float sx = (float)W / (float)(W - 1), sy = (float)H / (float)(H - 1);
    const float xmax = (float)(W + 1), ymax = (float)(H + 1);
    const float2 invV2 = make_float2(1.0f / (float)V, 1.0f / (float)V);
    const uint32_t bar = ptx::smem_u32(s_bar);

    for (int i = tid; i < nsrc * 3; i += blockDim.x) {
        const float *r = rt + (size_t)(b * nsrc + i / 3) * 12;
        const int k = i % 3;
        s_rt[i] = make_float4(r[3 * k], r[3 * k + 1], r[3 * k + 2], r[9 + k]);
    }
    for (int i = tid; i < d_end - d_begin; i += blockDim.x) s_dep[i] = __ldg(depth_values + (size_t)b * D + d_begin + i);
    if (tid == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&tmap);
    }
    uint32_t phase = 0;

    // reference-view texels of this pixel (view 0 of the same fp16 tensor): 4 chunks x 16 B
    const uint4 *ref_px = tex + (((size_t)b * V * H + min(y, H - 1)) * 4) * W + min(x, W - 1);

    int ds = d_begin;
    while (ds < d_end) {
        __syncthreads();  // previous segment's windows are no longer read; s_rt / s_dep visible (first pass)
        if (warp == 0) {
            // ---- plan the segment: the longest run of planes starting at ds whose footprint fits every window
            const float cx = (lane & 1) ? (float)min(tx0 + TW - 1, W - 1) : (float)tx0;
            const float cy = (lane & 2) ? (float)min(ty0 + TH - 1, H - 1) : (float)ty0;
            int L = d_end - ds;
            int my_mode = 0;
            int2 my_org = make_int2(0, 0);
            while (true) {
                float dlo = 3.0e38f, dhi = -3.0e38f;
                for (int i = lane; i < L; i += 32) {
                    const float dv = s_dep[ds - d_begin + i];
                    dlo = fminf(dlo, dv);
                    dhi = fmaxf(dhi, dv);
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    dlo = fminf(dlo, __shfl_xor_sync(0xffffffffu, dlo, o));
                    dhi = fmaxf(dhi, __shfl_xor_sync(0xffffffffu, dhi, o));
                }
                const bool dep_ok = (dlo > 0.f) && (dhi < 3.0e38f);  // also false for NaN depths
                const float cd = (lane & 4) ? dhi : dlo;
                bool allfit = true;
                for (int v = 0; v < nwin; ++v) {
                    const float4 c0 = s_rt[3 * v], c1 = s_rt[3 * v + 1], c2 = s_rt[3 * v + 2];
                    const float qz = fmaf(fmaf(c2.x, cx, fmaf(c2.y, cy, c2.z)), cd, c2.w);
                    const Coord c = project(c0, c1, c2, cx, cy, cd, sx, sy);
                    bool ok = dep_ok && (qz > 1e-20f) && (fabsf(c.ix) < 1.0e8f) && (fabsf(c.iy) < 1.0e8f);
                    float x_lo = ok ? c.ix : 0.f, x_hi = x_lo, y_lo = ok ? c.iy : 0.f, y_hi = y_lo;
#pragma unroll
                    for (int o = 4; o; o >>= 1) {
                        x_lo = fminf(x_lo, __shfl_xor_sync(0xffffffffu, x_lo, o));
                        x_hi = fmaxf(x_hi, __shfl_xor_sync(0xffffffffu, x_hi, o));
                        y_lo = fminf(y_lo, __shfl_xor_sync(0xffffffffu, y_lo, o));
                        y_hi = fmaxf(y_hi, __shfl_xor_sync(0xffffffffu, y_hi, o));
                    }
                    ok = __all_sync(0xffffffffu, ok);
                    // needed texel columns / rows, clipped to the part that can be non-zero: [-1, W] x [-1, H]
                    const int ex0 = max((int)floorf(x_lo - 0.02f), -1), ex1 = min((int)floorf(x_hi + 0.02f) + 1, W);
                    const int ey0 = max((int)floorf(y_lo - 0.02f), -1), ey1 = min((int)floorf(y_hi + 0.02f) + 1, H);
                    const bool fit = ok && (ex1 - ex0 < WX) && (ey1 - ey0 < WY);
                    if (lane == v) {
                        my_mode = fit ? 0 : 1;
                        // an empty clipped range (footprint entirely outside) gives ex1 < ex0: any origin works, every
                        // sample then fails the in-window test or reads zero fill
                        my_org = make_int2(max(min(ex0, W), -WX), max(min(ey0, H), -WY));
                    }
                    allfit &= fit;
                }
                if (allfit || L == 1) break;
                L = (L + 1) >> 1;
            }
            // nwin <= 32: lane v holds view v's decision
            if (lane < nwin) {
                s_org[lane] = my_org;
                s_mode[lane] = my_mode;
            }
            const unsigned loadmask = __ballot_sync(0xffffffffu, (lane < nwin) && (my_mode == 0));
            if (lane == 0) {
                s_seg[0] = L;
                ptx::mbar_arrive_expect_tx(bar, (uint32_t)__popc(loadmask) * win_bytes);
            }
            if ((lane < nwin) && (my_mode == 0))
                ptx::tma_load_4d(ptx::smem_u32(smem_raw + (size_t)lane * win_bytes), &tmap, bar, 2 * my_org.x, 0, my_org.y,
                                 b * V + 1 + lane);
        }
        __syncthreads();
        const int L = s_seg[0];
        ptx::mbar_wait(bar, phase);
        phase ^= 1;

        for (int d = ds; d < ds + L; ++d) {
            const float dep = s_dep[d - d_begin];
            float2 S[HACC ? 1 : 16], Q[HACC ? 1 : 16];
            __half2 Sh[HACC ? 16 : 1], Qh[HACC ? 16 : 1];
            uint4 nref[HACC ? 4 : 1];  // -x_ref, the start value of every interpolation chain
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 rv = __ldg(ref_px + (size_t)c * W);
                if constexpr (HACC) {
                    nref[c] = make_uint4(rv.x ^ 0x80008000u, rv.y ^ 0x80008000u, rv.z ^ 0x80008000u, rv.w ^ 0x80008000u);
#pragma unroll
                    for (int j = 0; j < 4; ++j) Sh[4 * c + j] = Qh[4 * c + j] = __float2half2_rn(0.f);
                } else {
                    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __half22float2(as_half2(rw[j]));
                        S[4 * c + j] = f;
                        Q[4 * c + j] = __fmul2_rn(f, f);
                    }
                }
            }
            for (int v = 0; v < nsrc; ++v) {
                const float4 c0 = s_rt[3 * v], c1 = s_rt[3 * v + 1], c2 = s_rt[3 * v + 2];
                Coord p = project(c0, c1, c2, xf, yf, dep, sx, sy);
                // non-finite -> far outside (CUDA grid_sampler rule); everything beyond one texel outside is zero anyway
                p.ix = fminf(fmaxf(p.ix, -2.f), xmax);
                p.iy = fminf(fmaxf(p.iy, -2.f), ymax);
                const float fx = floorf(p.ix), fy = floorf(p.iy);
                const float bx = p.ix - fx, by = p.iy - fy;
                const int x0 = (int)fx, y0 = (int)fy;
                float w11 = bx * by;
                float w10 = by - w11, w01 = bx - w11;
                float w00 = (1.0f - bx) - w10;
                const bool windowed = (v < nwin) && (s_mode[v] == 0);  // CTA-uniform
                if (windowed) {
                    const int2 org = s_org[v];
                    int ox = x0 - org.x, oy = y0 - org.y;
                    const bool in = ((unsigned)ox <= (unsigned)(WX - 2)) & ((unsigned)oy <= (unsigned)(WY - 2));
                    if (!in) {
                        ox = 0; oy = 0;
                        w00 = 0.f; w01 = 0.f; w10 = 0.f; w11 = 0.f;
                    }
                    const __half2 h00 = __float2half2_rn(w00), h01 = __float2half2_rn(w01);
                    const __half2 h10 = __float2half2_rn(w10), h11 = __float2half2_rn(w11);
                    const uint4 *wp = reinterpret_cast<const uint4 *>(smem_raw + (size_t)v * win_bytes) + oy * ROWQ + ox;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 ta = wp[c * WX], tb = wp[c * WX + 1], tc = wp[ROWQ + c * WX], td = wp[ROWQ + c * WX + 1];
                        if constexpr (HACC) accumulate_chunk_dev(ta, tb, tc, td, h00, h01, h10, h11, nref[c], Sh + 4 * c, Qh + 4 * c);
                        else accumulate_chunk(ta, tb, tc, td, h00, h01, h10, h11, S + 4 * c, Q + 4 * c);
                    }
                } else {
                    // per-tap global gather with explicit zero padding (footprint too large for a window)
                    const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
                    const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
                    const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
                    const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
                    const __half2 h00 = __float2half2_rn((vx0 & vy0) ? w00 : 0.f), h01 = __float2half2_rn((vx1 & vy0) ? w01 : 0.f);
                    const __half2 h10 = __float2half2_rn((vx0 & vy1) ? w10 : 0.f), h11 = __float2half2_rn((vx1 & vy1) ? w11 : 0.f);
                    const uint4 *img = tex + ((size_t)(b * V + 1 + v) * H) * 4 * W;
                    const uint4 *r0 = img + (size_t)cy0 * 4 * W, *r1 = img + (size_t)cy1 * 4 * W;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 ta = __ldg(r0 + c * W + cx0), tb = __ldg(r0 + c * W + cx1);
                        const uint4 tc = __ldg(r1 + c * W + cx0), td = __ldg(r1 + c * W + cx1);
                        if constexpr (HACC) accumulate_chunk_dev(ta, tb, tc, td, h00, h01, h10, h11, nref[c], Sh + 4 * c, Qh + 4 * c);
                        else accumulate_chunk(ta, tb, tc, td, h00, h01, h10, h11, S + 4 * c, Q + 4 * c);
                    }
                }
            }
            if (live) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // Q/V - (S/V)^2   (mvsnet.py:177), packed fp32x2
                        const float2 sv = HACC ? __half22float2(Sh[HACC ? 4 * c + j : 0]) : S[HACC ? 0 : 4 * c + j];
                        const float2 qv = HACC ? __half22float2(Qh[HACC ? 4 * c + j : 0]) : Q[HACC ? 0 : 4 * c + j];
                        const float2 m = __fmul2_rn(sv, invV2);
                        const float2 r = __ffma2_rn(qv, invV2, __fmul2_rn(m, make_float2(-m.x, -m.y)));
                        if constexpr (kActF16) {  // saturating: a variance above 65504 (|feature| > 250) stays finite
                            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(r.y), "f"(r.x));
                        } else {
                            const __nv_bfloat162 o = __floats2bfloat162_rn(r.x, r.y);
                            pk[j] = *reinterpret_cast<const uint32_t *>(&o);
                        }
                    }
                    __stcs(out + ((((size_t)b * 4 + c) * D + d) * H + y) * W + x, make_uint4(pk[0], pk[1], pk[2], pk[3]));
                }
            }
        }
        ds += L;
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
namespace {

struct WinPlan {
    int nwin;  // source views with a shared-memory window (the rest gather from global memory)
    int wy;    // rows per window
    size_t smem;
};

// rows per window: the tile's TH+1 rows plus room for vertical drift, within the per-CTA smem budget
WinPlan plan_windows(int nsrc, int th, int wx, int dchunk, int smem_budget) {
    const int rowb = 4 * wx * 16;
    const int misc = nsrc * 48 + 32 * 12 + dchunk * 4 + 64;
    WinPlan p;
    p.nwin = std::min(nsrc, 32);
    p.wy = 2;
    while (p.nwin > 0) {
        p.wy = std::min((smem_budget - misc) / (p.nwin * rowb), th + 8);
        if (p.wy >= th + 2) break;
        --p.nwin;  // more source views than fit: the last ones take the global-gather path
    }
    if (p.nwin == 0) p.wy = 2;
    p.smem = (size_t)p.nwin * p.wy * rowb + misc;
    return p;
}

int encode_window_map(CUtensorMap *tmap, const void *tex, int N, int H, int W, int WX, int WY) {
    tmap_encode_fn enc = get_tmap_encode();
    MVS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    // dims (x as uint64 pairs, chunk, y, n): the box lands in smem as [row][chunk][col][16 B]
    cuuint64_t gdim[4] = {(cuuint64_t)2 * W, 4, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)W * 64, (cuuint64_t)H * W * 64};
    cuuint32_t box[4] = {(cuuint32_t)(2 * WX), 4, (cuuint32_t)WY, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void *>(tex), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return set_error(MVS_ERR_CUDA, "cuTensorMapEncodeTiled (feature windows) failed (%d)", (int)cr);
    return MVS_OK;
}

template <int TWW, int TH, int WX, bool HACC>
int launch_win(const void *tex16, const float *rt, const float *depth_values, void *vol_cp8, int B, int V, int D, int H,
               int W, int dchunk, int smem_budget, cudaStream_t st) {
    const int nsrc = V - 1;
    const WinPlan p = plan_windows(nsrc, TH, WX, dchunk, smem_budget);
    CUtensorMap tmap;
    if (int rc = encode_window_map(&tmap, tex16, B * V, H, W, WX, p.wy)) return rc;
    auto kern = warp_variance_win_kernel<TWW, TH, WX, HACC>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MVS_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_budget));
        configured_dev = dev;
    }
    dim3 grid(cdiv(W, 32 * TWW), cdiv(H, TH), B * cdiv(D, dchunk));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(32 * TWW * TH);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MVS_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, (const uint4 *)tex16, (const float *)rt, (const float *)depth_values, (uint4 *)vol_cp8, V,
                                nsrc, p.nwin, D, H, W, dchunk, p.wy));
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

}  // namespace

// tex16: fp16 RCP8 features of all views [B*V][H][4][W][8].  Asynchronous on st.
// half_sums: packed-half sums of deviations from the reference view (faster; relative error of large variances up to
// 2^-6) instead of fp32 sums of the warped values (2^-7 everywhere, the tolerance stated for the tensor-core mode).
int warp_variance_windows(const void *tex16, const float *rt, const float *depth_values, void *vol_cp8, int B, int V, int D,
                          int H, int W, int half_sums, cudaStream_t st) {
    static const int cfg = [] {
        const char *e = getenv("MVS_WIN_CONFIG");  // tuning knob: 0 = 32x8 tile, 2 CTAs/SM; 1 = 64x8 tile, 1 CTA/SM
        return e ? atoi(e) : 0;
    }();
    int dchunk = 16;
    if (const char *e = getenv("MVS_WARP_DCHUNK")) dchunk = std::max(1, std::min(32, atoi(e)));
    while ((long long)B * cdiv(D, dchunk) > 65535 && dchunk < 32) dchunk <<= 1;
    MVS_REQUIRE((long long)B * cdiv(D, dchunk) <= 65535, "B*D=%lld too large for one launch", (long long)B * D);
    // MVS_WIN_HACC=0/1 overrides the caller's choice (tools/win_tune.py)
    static const int force = [] { const char *e = getenv("MVS_WIN_HACC"); return e ? atoi(e) : -1; }();
    const bool hacc = force >= 0 ? force != 0 : half_sums != 0;
    if (cfg == 1) return launch_win<2, 8, 80, false>(tex16, rt, depth_values, vol_cp8, B, V, D, H, W, dchunk, 227 * 1024, st);
    if (!hacc) return launch_win<1, 8, 40, false>(tex16, rt, depth_values, vol_cp8, B, V, D, H, W, dchunk, 113 * 1024, st);
    return launch_win<1, 8, 40, true>(tex16, rt, depth_values, vol_cp8, B, V, D, H, W, dchunk, 113 * 1024, st);
}

int features_nchw_to_rcp8(const float *fea, void *tex16, int N, int H, int W, cudaStream_t st) {
    const long long total = (long long)N * H * 4 * W;
    nchw_to_rcp8_kernel<<<cdiv(total, 256), 256, 0, st>>>(fea, (uint4 *)tex16, H, W, total);
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

int features_nhwc16_to_rcp8(const void *fea16, void *tex16, int N, int H, int W, cudaStream_t st) {
    const long long total = (long long)N * H * 4 * W;
    nhwc16_to_rcp8_kernel<<<cdiv(total, 256), 256, 0, st>>>((const uint4 *)fea16, (uint4 *)tex16, W, total);
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

}  // namespace mvs
