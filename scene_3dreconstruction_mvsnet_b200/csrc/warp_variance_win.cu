// Fused plane-sweep warp + variance, fourth generation: TMA-staged source WINDOWS in shared memory, 16x8 pixel tiles,
// window shape chosen per segment, packed-half deviation sums, fp16 volume out.
//
// Replaces, for the reference (olivier-2018/scene_3Dreconstruction_MVSNet), in the tensor-core precision mode:
//   models/module.py:96-139   homo_warping  (grid construction + F.grid_sample, bilinear, zero padding)
//   models/mvsnet.py:145-177  running sum / sum of squares over views and the variance
//
// Structure (what survives from generation 3):
//   * Features of ALL views are fp16 "RCP8" [n][y][chunk 0..3][x][8 ch]: a TMA box over (x, chunk, y) lands in shared
//     memory as [row][chunk][col][16 B]; TMA zero-fills outside the image, which IS grid_sample's padding_mode='zeros'.
//   * A CTA owns a pixel tile and a run of depth planes.  For a "segment" of planes one warp bounds the tile's footprint
//     in every source view from the 8 corners of (tile x depth range) -- the map (x*d, y*d, d) -> (u, v) is projective, so
//     the corners' bounding box contains every sample while q_z > 0 -- and issues one TMA box per view.
//   * Thread = pixel x 32 channels; 16 LDS.128 tap loads per (pixel, view, plane) from ONE address.
// What changed (round 2; measured reasons in profiles/r02_warp_kernel.md):
//   * Tile 16 x 8 instead of 32 x 8, two plane phases per CTA (threads 0-127 take the even planes of the segment, 128-255
//     the odd ones, same windows).  The footprint of a 32-wide tile under a 10 degree roll is 16 rows -- it never fitted the
//     11-row windows, every plane fell back to per-tap global gathers and DTU-like (rotated) cameras ran 2.6x slower than
//     rectified ones (2.84 vs 1.08 ms).  A 16-wide tile's footprint is 28 x 13 at 10 degrees.
//   * Three window shapes of the same size (wide / medium / tall), one tensor map each; the planner picks, per segment, the
//     first shape that holds the footprint of every view for the longest run of planes.
//   * Projection hoisted: per (pixel, view) a = R.(x, y, 1) lives in registers, a plane costs q = a*d + t (3 FMA), one
//     reciprocal and 2 FMA that land directly in window coordinates; clamping to the window replaces the bounds test.
//   * Running sums over the DEVIATION from the reference view (the variance is shift invariant) in packed half: the
//     interpolation chain starts from -x_ref, so the subtraction is free, and Sum d / Sum d^2 cost 2 HFMA2 per channel pair
//     and view.  The variance itself is formed in packed half as well and IS the fp16 volume element: no conversions.
//     On B200 HFMA2 issues at 0.5 / clk / scheduler and shares its pipe with every fp32 FMA (tools/pipe_microbench.cu),
//     so the 96 packed-half operations per (pixel, view, plane) -- 64 interpolation, 32 sums -- are one floor of this
//     formulation (0.46 ms at the DTU shape), the 256 bytes of LDS per (pixel, view, plane) the other (0.61 ms).
// Generality: a segment whose footprint fits no shape is halved; a single plane that still does not fit (extreme zoom,
// q_z <= 0 inside the tile, depth <= 0), or more source views than windows, takes a per-tap global gather with explicit
// clamps -- slow, exact.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace mvs {

namespace {

static_assert(kActF16, "the fused warp kernel writes the variance as packed half: the 16-bit volume format must be fp16");

constexpr int kC = 32;
constexpr int TW = 16, TH = 8;         // pixel tile
constexpr int kPhase = 2;              // plane phases per CTA
constexpr int kThreads = TW * TH * kPhase;
constexpr int kMaxWin = 8;             // source views that can have a window
constexpr int kShapes = 3;
constexpr int kMaxSeg = 32;            // planes per segment (and per CTA)

struct WinShapes {
    int wx[kShapes], wy[kShapes];
};

// Image index of every (batch element, view) in the feature tensor; n = 0: the dense form b * V + v.  The indexed form
// serves a feature POOL shared by the reference views of a scan (every image goes through FeatureNet once per scan).
constexpr int kMaxViewIds = 32;
struct ViewIds {
    int n;
    int id[kMaxViewIds];
};

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// shared-memory loads the compiler must leave where they are written (hoisting the per-view constants of all views out
// of the plane loop costs 32 registers and spills the accumulators)
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ __half2 as_half2(uint32_t u) { return *reinterpret_cast<const __half2 *>(&u); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<const uint32_t *>(&h); }

// one 8-channel chunk of one (pixel, view): 4 taps, interpolation chain seeded with -x_ref, deviation sums.
// FIRST: the first source view of a plane initialises the sums (no zeroing, no adds).
template <bool FIRST>
__device__ __forceinline__ void accumulate_chunk(const uint4 a, const uint4 b, const uint4 c, const uint4 d, const __half2 h00,
                                                 const __half2 h01, const __half2 h10, const __half2 h11, const uint4 nref,
                                                 __half2 *S, __half2 *Q) {
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
    const uint32_t wc[4] = {c.x, c.y, c.z, c.w}, wd[4] = {d.x, d.y, d.z, d.w};
    const uint32_t wr[4] = {nref.x, nref.y, nref.z, nref.w};  // -x_ref
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2 dv = __hfma2(as_half2(wd[j]), h11,
                                   __hfma2(as_half2(wc[j]), h10, __hfma2(as_half2(wb[j]), h01, __hfma2(as_half2(wa[j]), h00, as_half2(wr[j])))));
        if constexpr (FIRST) {
            S[j] = dv;
            Q[j] = __hmul2(dv, dv);
        } else {
            S[j] = __hadd2(S[j], dv);
            Q[j] = __hfma2(dv, dv, Q[j]);
        }
    }
}

// a source view whose sample lies outside its image contributes x_v = 0, i.e. the deviation -x_ref: exactly what the
// interpolation chain gives on four zero taps (0 * w + -x_ref), without the loads and the chain
template <bool FIRST>
__device__ __forceinline__ void accumulate_empty(const uint4 *nref, __half2 *S, __half2 *Q) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t wr[4] = {nref[c].x, nref[c].y, nref[c].z, nref[c].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __half2 dv = as_half2(wr[j]);
            if constexpr (FIRST) {
                S[4 * c + j] = dv;
                Q[4 * c + j] = __hmul2(dv, dv);
            } else {
                S[4 * c + j] = __hadd2(S[4 * c + j], dv);
                Q[4 * c + j] = __hfma2(dv, dv, Q[4 * c + j]);
            }
        }
    }
}

// bilinear weights of (bx, by) in [0, 1): fp32 products, rounded once to half, each broadcast to both halves
__device__ __forceinline__ void tap_weights(float bx, float by, __half2 &h00, __half2 &h01, __half2 &h10, __half2 &h11) {
    const float w11 = bx * by;
    const float w10 = by - w11, w01 = bx - w11;
    const float w00 = (1.0f - bx) - w10;
    h00 = __float2half2_rn(w00);
    h01 = __float2half2_rn(w01);
    h10 = __float2half2_rn(w10);
    h11 = __float2half2_rn(w11);
}

// variance of the V views from the deviation sums, packed half, saturated: (Q - S*(S/V)) / V     (mvsnet.py:177)
__device__ __forceinline__ uint32_t variance_f16x2(__half2 S, __half2 Q, __half2 invV) {
    const __half2 t = __hmul2(S, invV);
    const __half2 u = __hfma2(__hneg2(t), S, Q);   // the cancellation happens inside one fused operation
    __half2 r = __hmul2(u, invV);
    // inf / NaN (|deviation| > 255: sum of squares beyond fp16) -> largest finite value; -0 / tiny negatives -> 0
    r = __hmin2(r, __float2half2_rn(65504.f));
    r = __hmax2(r, __float2half2_rn(0.f));
    return as_u32(r);
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
        w[j] = as_u32(__floats2half2_rn(lo, hi));
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

// per segment and source view, written by the planner warp
struct SegView {
    float cx, cy;    // window coordinate = q.xy / q.z (scaled) + c:  c = -0.5 - window origin
    float hix, hiy;  // samples are clamped to [0, hi]: the planner guarantees every in-image sample is inside already
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// grid = (ceil(W / 16), ceil(H / 8), B * ceil(D / dchunk)), block = 256 threads:
//   thread t: plane phase t >> 7; a warp = a 16 x 2 pixel strip of the tile, lanes laid out per CTA (see "Lane -> pixel")
// dynamic smem: nwin windows of win_bytes | per-view constants | segment table | empty masks | depths | mbarrier
// NSRC > 0: number of source views known at compile time (view loop unrolled, a = R.(x,y,1) in registers);
// NSRC = 0: any number of views (a recomputed from shared memory per plane).
// ------------------------------------------------------------------------------------------------
template <int NSRC>
__global__ void __launch_bounds__(kThreads, 2)
warp_variance_win_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                         const __grid_constant__ CUtensorMap tmap2,
                         const uint4 *__restrict__ tex,           // fp16 RCP8 features, all views (the maps' memory)
                         const float *__restrict__ rt,            // [B*nsrc][12] rot(9) | trans(3)
                         const float *__restrict__ depth_values,  // [B,D]
                         uint4 *__restrict__ out,                 // fp16 CP8 [B,4,D,H,W,8]
                         int V, int nsrc_rt, int nwin, int D, int H, int W, int dchunk, const WinShapes shp, uint32_t win_bytes,
                         const ViewIds vid) {
    // the next kernel (conv0, launched with programmatic stream serialization) may start its set-up on SMs this grid has left
    ptx::pdl_launch_dependents();
    ptx::pdl_wait();  // homographies and features are read below
    const int nsrc = NSRC > 0 ? NSRC : nsrc_rt;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *sp = smem_raw + (size_t)nwin * win_bytes;
    float4 *s_rt = reinterpret_cast<float4 *>(sp);  // [nsrc][3]: (r0 r1 r2 t) per output coordinate, x / y rows pre-scaled
    sp += (size_t)nsrc * 48;
    SegView *s_sv = reinterpret_cast<SegView *>(sp);  // [nsrc]
    sp += (size_t)nsrc * 16;
    float4 *s_tv = reinterpret_cast<float4 *>(sp);  // [nsrc]: translation (x / y scaled) of each view
    sp += (size_t)nsrc * 16;
    uint32_t *s_emp = reinterpret_cast<uint32_t *>(sp);  // [nsrc]: bit i = plane ds + i of the segment is empty in view v
    sp += (size_t)nsrc * 4;
    float *s_dep = reinterpret_cast<float *>(sp);  // [dchunk]
    sp += (size_t)dchunk * 4;
    sp = smem_raw + (((size_t)(sp - smem_raw) + 15) & ~(size_t)15);  // (an integer round trip would make every access generic)
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(sp);
    int *s_seg = reinterpret_cast<int *>(sp + 8);  // [0] planes in the current segment, [1] window shape or -1 (gather)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int phase = tid >> 7;
    const int nchunks = (D + dchunk - 1) / dchunk;
    const int b = blockIdx.z / nchunks;
    const int d_begin = (blockIdx.z % nchunks) * dchunk;
    const int d_end = min(D, d_begin + dchunk);
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int img0 = b * V;  // index into vid.id / image index of the reference view in the dense form
    // ix = px * W/(W-1) - 0.5 is module.py:130-131 composed with grid_sample's align_corners=False un-normalisation
    const float sx = (float)W / (float)(W - 1), sy = (float)H / (float)(H - 1);
    const uint32_t bar = ptx::smem_u32(s_bar);
    const uint32_t win0 = ptx::smem_u32(smem_raw);
    const uint32_t sv0 = ptx::smem_u32(s_sv), tv0 = ptx::smem_u32(s_tv);

    for (int i = tid; i < nsrc * 3; i += kThreads) {
        const float *r = rt + (size_t)(b * nsrc + i / 3) * 12;
        const int k = i % 3;
        const float s = k == 0 ? sx : (k == 1 ? sy : 1.0f);
        s_rt[i] = make_float4(r[3 * k] * s, r[3 * k + 1] * s, r[3 * k + 2] * s, r[9 + k] * s);
    }
    for (int i = tid; i < nsrc; i += kThreads) {
        const float *r = rt + (size_t)(b * nsrc + i) * 12;
        s_tv[i] = make_float4(r[9] * sx, r[10] * sy, r[11], 0.f);
    }
    for (int i = tid; i < d_end - d_begin; i += kThreads) s_dep[i] = __ldg(depth_values + (size_t)b * D + d_begin + i);
    if (tid == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&tmap0);
        ptx::prefetch_tensormap(&tmap1);
        ptx::prefetch_tensormap(&tmap2);
    }
    uint32_t bar_phase = 0;
    __syncthreads();  // s_rt, s_dep

    // Lane -> pixel mapping, chosen per CTA.  A warp always covers a 16 x 2 pixel strip; the 8 lanes of a quarter warp
    // (one LDS.128 wavefront) take either 8 consecutive pixels of one row -- conflict-free while their samples stay in
    // one window row and advance by at most one texel per pixel: rectified cameras -- or a 4 x 2 block: the window row
    // pitch is 64 mod 128 bytes (odd number of columns), so the block's two pixel rows land in disjoint halves of the 8
    // bank groups, which survives roll and zoom far better (rotated cameras: 34-76 % excess wavefronts with the row
    // mapping, 23 % with blocks; rectified: the row mapping is 3 % faster, ncu r2k).
    if (warp == 0) {
        // lane = (pixel 4 or 11 of the tile's middle row, first / last plane, view slot): the displacement between the two
        // pixels, 7 apart, in every view at both ends of the depth range (more than 8 views: the block mapping, always safe)
        const int e = lane & 1, k = (lane >> 1) & 1, v = lane >> 2;
        const float cyp = (float)min(ty0 + TH / 2, H - 1);
        const float xx = (float)min(tx0 + 4, W - 1) + 7.f * e;
        const float dep = s_dep[k ? d_end - d_begin - 1 : 0];
        const int vv = v < nsrc ? v : 0;
        const float4 c0 = s_rt[3 * vv], c1 = s_rt[3 * vv + 1], c2 = s_rt[3 * vv + 2];
        const float iz = rcp_approx(fmaf(fmaf(c2.x, xx, fmaf(c2.y, cyp, c2.z)), dep, c2.w));
        const float px = fmaf(fmaf(c0.x, xx, fmaf(c0.y, cyp, c0.z)), dep, c0.w) * iz;
        const float py = fmaf(fmaf(c1.x, xx, fmaf(c1.y, cyp, c1.z)), dep, c1.w) * iz;
        const float dx = __shfl_xor_sync(0xffffffffu, px, 1) - px, dy = __shfl_xor_sync(0xffffffffu, py, 1) - py;
        const bool ok = (e != 0) || (v >= nsrc) || ((fabsf(dy) < 0.25f) && (dx > 5.0f) && (dx < 7.15f));  // false for NaN
        const bool rows_ok = __all_sync(0xffffffffu, ok) && nsrc <= 8;
        if (lane == 0) s_seg[2] = rows_ok ? 0 : 1;
    }
    __syncthreads();
    const bool blk42 = s_seg[2] != 0;
    const int x = blk42 ? tx0 + 4 * ((tid >> 3) & 3) + (tid & 3) : tx0 + (tid & 15);
    const int y = blk42 ? ty0 + 2 * ((tid >> 5) & 3) + ((tid >> 2) & 1) : ty0 + ((tid >> 4) & 7);
    const bool live = (x < W) & (y < H);
    const float xf = (float)x, yf = (float)y;

    // -x_ref of this pixel (view 0 of the same fp16 tensor), the seed of every interpolation chain: 4 chunks x 16 B
    uint4 nref[4];
    {
        const int ref_img = vid.n ? vid.id[img0] : img0;
        const uint4 *ref_px = tex + (((size_t)ref_img * H + min(y, H - 1)) * 4) * W + min(x, W - 1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 rv = __ldg(ref_px + (size_t)c * W);
            nref[c] = make_uint4(rv.x ^ 0x80008000u, rv.y ^ 0x80008000u, rv.z ^ 0x80008000u, rv.w ^ 0x80008000u);
        }
    }
    const __half2 invV = __float2half2_rn(1.0f / (float)V);
    const size_t hw = (size_t)H * W, dhw = (size_t)D * hw;
    uint4 *const out_px = out + (size_t)b * 4 * dhw + (size_t)min(y, H - 1) * W + min(x, W - 1);  // + c * dhw + d * hw

    // a = R.(x, y, 1) per view (x / y rows already scaled): registers when the view count is a template parameter
    float ax[NSRC > 0 ? NSRC : 1], ay[NSRC > 0 ? NSRC : 1], az[NSRC > 0 ? NSRC : 1];
    if constexpr (NSRC > 0) {
#pragma unroll
        for (int v = 0; v < NSRC; ++v) {
            const float4 c0 = s_rt[3 * v], c1 = s_rt[3 * v + 1], c2 = s_rt[3 * v + 2];
            ax[v] = fmaf(c0.x, xf, fmaf(c0.y, yf, c0.z));
            ay[v] = fmaf(c1.x, xf, fmaf(c1.y, yf, c1.z));
            az[v] = fmaf(c2.x, xf, fmaf(c2.y, yf, c2.z));
        }
    }

    int ds = d_begin;
    while (ds < d_end) {
        __syncthreads();  // the previous segment's windows and table are no longer read
#define WARP_PLAN_TEXEL_BYTES 64
#include "warp_window_plan.inc"
#undef WARP_PLAN_TEXEL_BYTES
        __syncthreads();
        const int L = s_seg[0], shape = s_seg[1];
        if (shape >= 0) {
            // ================= windowed segment (the fast path) =================
            const int wx = shp.wx[shape];
            const uint32_t chb = (uint32_t)wx * 16u;   // bytes between the 8-channel chunks of a window row
            const uint32_t rowb = chb * 4u;            // bytes per window row
            uint32_t emp[NSRC > 0 ? NSRC : 1];
            if constexpr (NSRC > 0) {
#pragma unroll
                for (int v = 0; v < NSRC; ++v) emp[v] = s_emp[v];
            }
            ptx::mbar_wait(bar, bar_phase);
            bar_phase ^= 1;
            for (int d = ds + phase; d < ds + L; d += kPhase) {
                const float dep = s_dep[d - d_begin];
                const uint32_t dbit = 1u << (d - ds);
                __half2 S[16], Q[16];
                if constexpr (NSRC == 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) S[j] = Q[j] = __float2half2_rn(0.f);
                }
#pragma unroll
                for (int v = 0; v < (NSRC > 0 ? NSRC : 1); ++v) {
                    for (int vr = (NSRC > 0 ? v : 0); vr < (NSRC > 0 ? v + 1 : nsrc); ++vr) {  // runtime view loop when NSRC = 0
                        if ((NSRC > 0 ? emp[v] : s_emp[vr]) & dbit) {  // CTA-uniform: nothing of this view under the tile
                            if (NSRC > 0 && v == 0) accumulate_empty<true>(nref, S, Q);
                            else accumulate_empty<false>(nref, S, Q);
                            continue;
                        }
                        const float4 tv = lds_f4(tv0 + 16u * vr);
                        const float4 sv = lds_f4(sv0 + 16u * vr);
                        float pax, pay, paz;
                        if constexpr (NSRC > 0) {
                            pax = ax[v]; pay = ay[v]; paz = az[v];
                        } else {
                            const float4 c0 = s_rt[3 * vr], c1 = s_rt[3 * vr + 1], c2 = s_rt[3 * vr + 2];
                            pax = fmaf(c0.x, xf, fmaf(c0.y, yf, c0.z));
                            pay = fmaf(c1.x, xf, fmaf(c1.y, yf, c1.z));
                            paz = fmaf(c2.x, xf, fmaf(c2.y, yf, c2.z));
                        }
                        const float iz = rcp_approx(fmaf(paz, dep, tv.z));
                        float fx = fmaf(fmaf(pax, dep, tv.x), iz, sv.x);
                        float fy = fmaf(fmaf(pay, dep, tv.y), iz, sv.y);
                        fx = fminf(fmaxf(fx, 0.f), sv.z);   // NaN -> 0: memory safe whatever the coordinates
                        fy = fminf(fmaxf(fy, 0.f), sv.w);
                        const int ox = (int)fx, oy = (int)fy;      // >= 0: truncation is floor
                        const float bx = fx - (float)ox, by = fy - (float)oy;
                        __half2 h00, h01, h10, h11;
                        tap_weights(bx, by, h00, h01, h10, h11);
                        const uint32_t a0 = win0 + (uint32_t)vr * win_bytes + (uint32_t)oy * rowb + (uint32_t)ox * 16u;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint32_t a = a0 + (uint32_t)c * chb;
                            const uint4 ta = lds_u4(a), tb = lds_u4(a + 16u), tc = lds_u4(a + rowb), td = lds_u4(a + rowb + 16u);
                            if (NSRC > 0 && v == 0) accumulate_chunk<true>(ta, tb, tc, td, h00, h01, h10, h11, nref[c], S + 4 * c, Q + 4 * c);
                            else accumulate_chunk<false>(ta, tb, tc, td, h00, h01, h10, h11, nref[c], S + 4 * c, Q + 4 * c);
                        }
                    }
                }
                if (live) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 pk;
                        pk.x = variance_f16x2(S[4 * c], Q[4 * c], invV);
                        pk.y = variance_f16x2(S[4 * c + 1], Q[4 * c + 1], invV);
                        pk.z = variance_f16x2(S[4 * c + 2], Q[4 * c + 2], invV);
                        pk.w = variance_f16x2(S[4 * c + 3], Q[4 * c + 3], invV);
                        __stcs(out_px + (size_t)d * hw + (size_t)c * dhw, pk);
                    }
                }
            }
        } else {
            // ================= gather segment: per-tap global loads with explicit zero padding (exact, slow) ==========
            const float xmax = (float)(W + 1), ymax = (float)(H + 1);
            for (int d = ds + phase; d < ds + L; d += kPhase) {
                const float dep = s_dep[d - d_begin];
                __half2 S[16], Q[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) S[j] = Q[j] = __float2half2_rn(0.f);
                for (int v = 0; v < nsrc; ++v) {
                    const float4 c0 = s_rt[3 * v], c1 = s_rt[3 * v + 1], c2 = s_rt[3 * v + 2];
                    const float iz = rcp_approx(fmaf(fmaf(c2.x, xf, fmaf(c2.y, yf, c2.z)), dep, c2.w));
                    float ix = fmaf(fmaf(fmaf(c0.x, xf, fmaf(c0.y, yf, c0.z)), dep, c0.w), iz, -0.5f);
                    float iy = fmaf(fmaf(fmaf(c1.x, xf, fmaf(c1.y, yf, c1.z)), dep, c1.w), iz, -0.5f);
                    // non-finite -> far outside (CUDA grid_sampler rule); everything beyond one texel outside is zero anyway
                    ix = fminf(fmaxf(ix, -2.f), xmax);
                    iy = fminf(fmaxf(iy, -2.f), ymax);
                    const float fx = floorf(ix), fy = floorf(iy);
                    const int x0 = (int)fx, y0 = (int)fy;
                    const float bx = ix - fx, by = iy - fy;
                    const float w11 = bx * by, w10 = by - w11, w01 = bx - w11, w00 = (1.0f - bx) - w10;
                    const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
                    const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
                    const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
                    const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
                    const __half2 h00 = __float2half2_rn((vx0 & vy0) ? w00 : 0.f), h01 = __float2half2_rn((vx1 & vy0) ? w01 : 0.f);
                    const __half2 h10 = __float2half2_rn((vx0 & vy1) ? w10 : 0.f), h11 = __float2half2_rn((vx1 & vy1) ? w11 : 0.f);
                    const uint4 *img = tex + ((size_t)(vid.n ? vid.id[img0 + 1 + v] : img0 + 1 + v) * H) * 4 * W;
                    const uint4 *r0 = img + (size_t)cy0 * 4 * W, *r1 = img + (size_t)cy1 * 4 * W;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 ta = __ldg(r0 + c * W + cx0), tb = __ldg(r0 + c * W + cx1);
                        const uint4 tc = __ldg(r1 + c * W + cx0), td = __ldg(r1 + c * W + cx1);
                        accumulate_chunk<false>(ta, tb, tc, td, h00, h01, h10, h11, nref[c], S + 4 * c, Q + 4 * c);
                    }
                }
                if (live) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 pk;
                        pk.x = variance_f16x2(S[4 * c], Q[4 * c], invV);
                        pk.y = variance_f16x2(S[4 * c + 1], Q[4 * c + 1], invV);
                        pk.z = variance_f16x2(S[4 * c + 2], Q[4 * c + 2], invV);
                        pk.w = variance_f16x2(S[4 * c + 3], Q[4 * c + 3], invV);
                        __stcs(out_px + (size_t)d * hw + (size_t)c * dhw, pk);
                    }
                }
            }
        }
        ds += L;
    }
}

// ================================================================================================
// Strict-fp32 form of the same kernel (precision = "fp32", the reference's arithmetic): fp32 texels in "RCP4"
// [n][y][chunk 0..7][x][4 ch] -- again 16 bytes per (chunk, texel), so windows, planner and tap addressing are those of the
// fp16 kernel with eight chunks per texel instead of four -- the sample position of every (pixel, view, plane) in the
// reference's exact fp32 operation order (IEEE divisions, no contraction: module.py:119-136 composed with grid_sample's
// un-normalisation, the arithmetic of warp_variance_fwd2_kernel in warp_variance.cu), zero padding through masked
// interpolation factors as there, running sum / sum of squares over the views in fp32, fp32 NCDHW volume out.  The planner
// only decides WHERE the windows go (approximate arithmetic, generous margins); what a thread computes from a window is
// bit-identical to what the per-tap global gathers of warp_variance_fwd2_kernel compute.
// One CTA of 512 threads per SM (windows of 128-byte texels need all of shared memory): thread = pixel x 16 channels,
// a warp = 16 pixels x 2 channel halves, which share the sample positions through shuffles (each half computes every other view).
// ================================================================================================
namespace {

constexpr int kThreads32 = 2 * kThreads;

__device__ __forceinline__ float safe_coord32(float v) {
    return (v <= 2147483520.0f && v >= -2147483648.0f) ? v : -100.0f;  // NaN fails both compares (CUDA grid_sampler rule)
}

// exact sample of pixel (rx, ry, rz) = R.(x, y, 1) at depth d: masked factors (ax, bx, ay, by) and the texel (x0, y0)
__device__ __forceinline__ void sample_exact(float rx, float ry, float rz, float tx, float ty, float tz, float d, int H, int W,
                                             float4 &f, int &x0, int &y0) {
    const float qx = __fadd_rn(__fmul_rn(rx, d), tx);
    const float qy = __fadd_rn(__fmul_rn(ry, d), ty);
    const float qz = __fadd_rn(__fmul_rn(rz, d), tz);
    const float px = __fdiv_rn(qx, qz);
    const float py = __fdiv_rn(qy, qz);
    const float gx = __fsub_rn(__fdiv_rn(px, (float)(W - 1) * 0.5f), 1.0f);
    const float gy = __fsub_rn(__fdiv_rn(py, (float)(H - 1) * 0.5f), 1.0f);
    const float ix = safe_coord32(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)W), 1.0f), 0.5f));
    const float iy = safe_coord32(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)H), 1.0f), 0.5f));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    x0 = (int)fx0;
    y0 = (int)fy0;
    const int x1 = x0 + 1, y1 = y0 + 1;
    const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
    const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
    f.x = vx0 ? __fsub_rn(__fadd_rn(fx0, 1.0f), ix) : 0.f;
    f.y = vx1 ? __fsub_rn(ix, fx0) : 0.f;
    f.z = vy0 ? __fsub_rn(__fadd_rn(fy0, 1.0f), iy) : 0.f;
    f.w = vy1 ? __fsub_rn(iy, fy0) : 0.f;
}

__device__ __forceinline__ void accumulate4(const float4 a, const float4 b, const float4 c, const float4 d, float w00, float w01,
                                            float w10, float w11, float *S, float *Q) {
    const float vx = fmaf(d.x, w11, fmaf(c.x, w10, fmaf(b.x, w01, a.x * w00)));
    const float vy = fmaf(d.y, w11, fmaf(c.y, w10, fmaf(b.y, w01, a.y * w00)));
    const float vz = fmaf(d.z, w11, fmaf(c.z, w10, fmaf(b.z, w01, a.z * w00)));
    const float vw = fmaf(d.w, w11, fmaf(c.w, w10, fmaf(b.w, w01, a.w * w00)));
    S[0] += vx; Q[0] = fmaf(vx, vx, Q[0]);
    S[1] += vy; Q[1] = fmaf(vy, vy, Q[1]);
    S[2] += vz; Q[2] = fmaf(vz, vz, Q[2]);
    S[3] += vw; Q[3] = fmaf(vw, vw, Q[3]);
}

// fp32 NCHW [N][32][HW] -> fp32 RCP4 [N][H][8][W][4]; one thread per output float4
__global__ void nchw_to_rcp4_kernel(const float *__restrict__ in, float4 *__restrict__ out, int H, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W);
    long long r = i / W;
    const int c = (int)(r & 7);
    r >>= 3;
    const int y = (int)(r % H);
    const long long n = r / H;
    const size_t HW = (size_t)H * W;
    const float *src = in + ((size_t)n * kC + 4 * c) * HW + (size_t)y * W + x;
    out[i] = make_float4(__ldg(src), __ldg(src + HW), __ldg(src + 2 * HW), __ldg(src + 3 * HW));
}

}  // namespace

template <int NSRC>
__global__ void __launch_bounds__(kThreads32, 1)
warp_variance_win32_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                           const __grid_constant__ CUtensorMap tmap2,
                           const float4 *__restrict__ tex,           // fp32 RCP4 features of all views (the maps' memory)
                           const float *__restrict__ rt,            // [B*nsrc][12] rot(9) | trans(3)
                           const float *__restrict__ depth_values,  // [B,D]
                           float *__restrict__ out,                 // fp32 [B,32,D,H,W]
                           int V, int nsrc_rt, int nwin, int D, int H, int W, int dchunk, const WinShapes shp, uint32_t win_bytes) {
    const int nsrc = NSRC > 0 ? NSRC : nsrc_rt;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *sp = smem_raw + (size_t)nwin * win_bytes;
    float4 *s_rt = reinterpret_cast<float4 *>(sp);  // [nsrc][3]: the planner's pre-scaled (r0 r1 r2 t) rows
    sp += (size_t)nsrc * 48;
    SegView *s_sv = reinterpret_cast<SegView *>(sp);  // [nsrc]
    sp += (size_t)nsrc * 16;
    float4 *s_raw = reinterpret_cast<float4 *>(sp);  // [nsrc][3]: rot | trans as given (exact arithmetic)
    sp += (size_t)nsrc * 48;
    uint32_t *s_emp = reinterpret_cast<uint32_t *>(sp);  // [nsrc]
    sp += (size_t)nsrc * 4;
    float *s_dep = reinterpret_cast<float *>(sp);  // [dchunk]
    sp += (size_t)dchunk * 4;
    sp = smem_raw + (((size_t)(sp - smem_raw) + 15) & ~(size_t)15);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(sp);
    int *s_seg = reinterpret_cast<int *>(sp + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // a warp = the 16 pixels of one tile row x the two channel halves (lane bit 4): the halves of a pixel share the sample
    // positions through shuffles; warps 0-7 take the even planes of a segment, 8-15 the odd ones
    const int phase = warp >> 3, half = lane >> 4;
    const int nchunks = (D + dchunk - 1) / dchunk;
    const int b = blockIdx.z / nchunks;
    const int d_begin = (blockIdx.z % nchunks) * dchunk;
    const int d_end = min(D, d_begin + dchunk);
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int img0 = b * V;
    const float sx = (float)W / (float)(W - 1), sy = (float)H / (float)(H - 1);
    const uint32_t bar = ptx::smem_u32(s_bar);
    const uint32_t win0 = ptx::smem_u32(smem_raw);
    const uint32_t sv0 = ptx::smem_u32(s_sv);
    ViewIds vid;
    vid.n = 0;  // dense images b * V + v

    for (int i = tid; i < nsrc * 3; i += kThreads32) {
        const float *r = rt + (size_t)(b * nsrc + i / 3) * 12;
        const int k = i % 3;
        const float s = k == 0 ? sx : (k == 1 ? sy : 1.0f);
        s_rt[i] = make_float4(r[3 * k] * s, r[3 * k + 1] * s, r[3 * k + 2] * s, r[9 + k] * s);
        s_raw[i] = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);  // the 12 floats, four at a time
    }
    for (int i = tid; i < d_end - d_begin; i += kThreads32) s_dep[i] = __ldg(depth_values + (size_t)b * D + d_begin + i);
    if (tid == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&tmap0);
        ptx::prefetch_tensormap(&tmap1);
        ptx::prefetch_tensormap(&tmap2);
    }
    uint32_t bar_phase = 0;
    __syncthreads();

    // lane -> pixel: rows of 16 pixels (the block mapping of the fp16 kernel relies on a row pitch of 64 mod 128 bytes,
    // which 128-byte texels do not give)
    const int x = tx0 + (lane & 15), y = ty0 + (warp & 7);
    const bool live = (x < W) & (y < H);
    const int xc = min(x, W - 1), yc = min(y, H - 1);
    const float xf = (float)xc, yf = (float)yc;
    const size_t hw = (size_t)H * W;
    const float invV = 1.0f / (float)V;

    // this thread's 16 channels of the reference view: chunks 4 * half .. 4 * half + 3
    float ref[16];
    {
        const float4 *rp = tex + (((size_t)img0 * H + yc) * 8 + 4 * half) * W + xc;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 v = __ldg(rp + (size_t)c * W);
            ref[4 * c] = v.x; ref[4 * c + 1] = v.y; ref[4 * c + 2] = v.z; ref[4 * c + 3] = v.w;
        }
    }
    float *const out_px = out + ((size_t)b * kC + 16 * half) * D * hw + (size_t)yc * W + xc;  // + ch * D*hw + d * hw

    // R.(x, y, 1) per view in the reference's operation order (registers when the view count is a template parameter)
    // The two channel halves of a pixel need the same sample positions: half h computes views h, h + 2, ... (slot s = view
    // 2 s + h; an odd view count repeats the last view in the spare slot) and every lane fetches a view's result from its
    // owner with six shuffles -- the exact projection (four IEEE divisions) is ~135 instructions, more than the 96 FMAs it
    // feeds in a 16-channel thread.
    constexpr int NS = NSRC > 0 ? (NSRC + 1) / 2 : 1;
    float rx[NS], ry[NS], rz[NS];
    auto rot_exact = [&](int v, float &ox, float &oy, float &oz) {
        const float *r = reinterpret_cast<const float *>(s_raw + 3 * v);
        ox = __fadd_rn(__fadd_rn(__fmul_rn(r[0], xf), __fmul_rn(r[1], yf)), r[2]);
        oy = __fadd_rn(__fadd_rn(__fmul_rn(r[3], xf), __fmul_rn(r[4], yf)), r[5]);
        oz = __fadd_rn(__fadd_rn(__fmul_rn(r[6], xf), __fmul_rn(r[7], yf)), r[8]);
    };
    if constexpr (NSRC > 0) {
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) rot_exact(min(2 * sl + half, NSRC - 1), rx[sl], ry[sl], rz[sl]);
    }

    auto store_plane = [&](int d, const float *S, const float *Q) {
        if (!live) return;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float m = S[j] * invV;
            __stcs(out_px + (size_t)j * D * hw + (size_t)d * hw, fmaf(Q[j], invV, -m * m));  // Q/V - (S/V)^2   (mvsnet.py:177)
        }
    };

    int ds = d_begin;
    while (ds < d_end) {
        __syncthreads();  // the previous segment's windows and table are no longer read
#define WARP_PLAN_TEXEL_BYTES 128
#include "warp_window_plan.inc"
#undef WARP_PLAN_TEXEL_BYTES
        __syncthreads();
        const int L = s_seg[0], shape = s_seg[1];
        const bool windowed = shape >= 0;
        const int wx = windowed ? shp.wx[shape] : 2;
        const uint32_t chb = (uint32_t)wx * 16u;   // bytes between the 4-channel chunks of a window row
        const uint32_t rowb = chb * 8u;            // bytes per window row
        if (windowed) {
            ptx::mbar_wait(bar, bar_phase);
            bar_phase ^= 1;
        }
        for (int d = ds + phase; d < ds + L; d += kPhase) {
            const float dep = s_dep[d - d_begin];
            const uint32_t dbit = 1u << (d - ds);
            float S[16], Q[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                S[j] = ref[j];
                Q[j] = ref[j] * ref[j];
            }
            // one view's contribution from its masked factors f and top-left texel (x0, y0)
            auto add_view = [&](int vr, const float4 f, int x0, int y0) {
                const float w00 = f.x * f.z, w01 = f.y * f.z, w10 = f.x * f.w, w11 = f.y * f.w;
                if (windowed) {
                    // window coordinates of the top-left tap; a tap the planner did not cover is outside the image and
                    // has a zero factor: any in-window address will do for it
                    const float4 sv = lds_f4(sv0 + 16u * vr);
                    const int ox = -(int)(sv.x + 0.5f), oy = -(int)(sv.y + 0.5f);   // sv.x = -0.5 - origin
                    const int col = min(max(min(max(x0, -2), W + 1) - ox, 0), (int)sv.z);
                    const int row = min(max(min(max(y0, -2), H + 1) - oy, 0), (int)sv.w);
                    const uint32_t a0 = win0 + (uint32_t)vr * win_bytes + (uint32_t)row * rowb + (uint32_t)col * 16u +
                                        (uint32_t)(4 * half) * chb;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t a = a0 + (uint32_t)c * chb;
                        const float4 ta = lds_f4(a), tb = lds_f4(a + 16u), tc = lds_f4(a + rowb), td = lds_f4(a + rowb + 16u);
                        accumulate4(ta, tb, tc, td, w00, w01, w10, w11, S + 4 * c, Q + 4 * c);
                    }
                } else {
                    // gather segment: per-tap global loads at clamped texels (masked factors make the padding)
                    const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);   // |x0| < 2^31 - 128: no overflow
                    const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
                    const float4 *img = tex + ((size_t)(img0 + 1 + vr) * H) * 8 * W + (size_t)(4 * half) * W;
                    const float4 *r0 = img + (size_t)cy0 * 8 * W, *r1 = img + (size_t)cy1 * 8 * W;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 ta = __ldg(r0 + c * W + cx0), tb = __ldg(r0 + c * W + cx1);
                        const float4 tc = __ldg(r1 + c * W + cx0), td = __ldg(r1 + c * W + cx1);
                        accumulate4(ta, tb, tc, td, w00, w01, w10, w11, S + 4 * c, Q + 4 * c);
                    }
                }
            };
            if constexpr (NSRC > 0) {
#pragma unroll
                for (int sl = 0; sl < NS; ++sl) {
                    const int k0 = 2 * sl, k1 = min(2 * sl + 1, NSRC - 1);
                    // a view with nothing under the tile in this plane (CTA-uniform) adds exact zeros to both sums
                    const bool e0 = (s_emp[k0] & dbit) != 0, e1 = (2 * sl + 1 >= NSRC) || (s_emp[k1] & dbit) != 0;
                    if (e0 && e1) continue;
                    const int kown = min(2 * sl + half, NSRC - 1);
                    const float *traw = reinterpret_cast<const float *>(s_raw + 3 * kown) + 9;
                    float4 f;
                    int x0, y0;
                    sample_exact(rx[sl], ry[sl], rz[sl], traw[0], traw[1], traw[2], dep, H, W, f, x0, y0);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (j ? e1 : e0) continue;
                        const int src = (lane & 15) | (j << 4);   // the lane of this pixel in the half that owns view 2 sl + j
                        float4 fk;
                        fk.x = __shfl_sync(0xffffffffu, f.x, src);
                        fk.y = __shfl_sync(0xffffffffu, f.y, src);
                        fk.z = __shfl_sync(0xffffffffu, f.z, src);
                        fk.w = __shfl_sync(0xffffffffu, f.w, src);
                        const int xk = __shfl_sync(0xffffffffu, x0, src), yk = __shfl_sync(0xffffffffu, y0, src);
                        add_view(2 * sl + j, fk, xk, yk);
                    }
                }
            } else {
                for (int vr = 0; vr < nsrc; ++vr) {  // any number of views: every lane computes every view
                    if (s_emp[vr] & dbit) continue;
                    float prx, pry, prz;
                    rot_exact(vr, prx, pry, prz);
                    const float *traw = reinterpret_cast<const float *>(s_raw + 3 * vr) + 9;
                    float4 f;
                    int x0, y0;
                    sample_exact(prx, pry, prz, traw[0], traw[1], traw[2], dep, H, W, f, x0, y0);
                    add_view(vr, f, x0, y0);
                }
            }
            store_plane(d, S, Q);
        }
        ds += L;
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kSmemBudget = 113 * 1024;  // per CTA: two CTAs per SM

struct WinPlan {
    int nwin;            // source views with a shared-memory window (0: every segment gathers from global memory)
    uint32_t win_bytes;  // bytes per window slot
    WinShapes shp;
    size_t smem;
};

// Three shapes with (nearly) the same number of texels: wide for rectified cameras and large disparity steps, tall for
// rolled cameras.  Every shape holds at least the tile plus the bilinear halo.
// texb: bytes per texel (64: fp16 RCP8, 128: fp32 RCP4); budget: shared memory per CTA; view_bytes: per-view constants
WinPlan plan_windows(int nsrc, int dchunk, int texb = 64, int budget = kSmemBudget, int view_bytes = 84) {
    WinPlan p = {};
    const int misc = nsrc * view_bytes + dchunk * 4 + 64;
    p.nwin = std::min(nsrc, kMaxWin);
    if (nsrc > kMaxWin) p.nwin = 0;
    int texels = p.nwin > 0 ? (budget - misc) / (p.nwin * texb) : 0;
    texels = std::min(texels, 1024);  // TMA box: <= 256 elements per dimension, keep windows sane for few views
    if (p.nwin > 0 && texels < (TW + 4) * (TH + 2)) p.nwin = 0;  // too many views for useful windows
    if (p.nwin == 0) {
        for (int s = 0; s < kShapes; ++s) p.shp.wx[s] = p.shp.wy[s] = 2;
        p.win_bytes = 0;
        p.smem = misc;
        return p;
    }
    const int rows[kShapes] = {TH + 2, TH + 6, TH + 10};
    for (int s = 0; s < kShapes; ++s) {
        int wy = rows[s];
        // odd number of columns: row pitch = wx * 64 B = 64 mod 128 (see the lane mapping in the kernel)
        int wx = (std::min(texels / wy, 121) - 1) | 1;
        if (wx < TW + 4) {  // few texels: fall back to the flattest shape that holds the tile
            wy = TH + 2;
            wx = (texels / wy - 1) | 1;
        }
        p.shp.wx[s] = wx;
        p.shp.wy[s] = wy;
    }
    p.win_bytes = (uint32_t)(((size_t)texels * texb + 127) / 128 * 128);
    while ((size_t)p.nwin * p.win_bytes + misc > (size_t)budget) p.win_bytes -= 128;
    for (int s = 0; s < kShapes; ++s)
        while ((uint32_t)(p.shp.wx[s] * p.shp.wy[s] * texb) > p.win_bytes) p.shp.wx[s] -= 2;
    p.smem = (size_t)p.nwin * p.win_bytes + misc;
    return p;
}

int encode_window_map(CUtensorMap *tmap, const void *tex, int N, int H, int W, int WX, int WY, int chunks = 4) {
    tmap_encode_fn enc = get_tmap_encode();
    MVS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    // dims (x as uint64 pairs, chunk, y, n): the box lands in smem as [row][chunk][col][16 B]
    cuuint64_t gdim[4] = {(cuuint64_t)2 * W, (cuuint64_t)chunks, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)W * 16 * chunks, (cuuint64_t)H * W * 16 * chunks};
    cuuint32_t box[4] = {(cuuint32_t)(2 * WX), (cuuint32_t)chunks, (cuuint32_t)WY, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void *>(tex), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return set_error(MVS_ERR_CUDA, "cuTensorMapEncodeTiled (feature windows) failed (%d)", (int)cr);
    return MVS_OK;
}

template <int NSRC>
int launch_win(const CUtensorMap *maps, const WinPlan &p, const void *tex16, const float *rt, const float *depth_values,
               void *vol_cp8, int B, int V, int D, int H, int W, int dchunk, const ViewIds &vid, cudaStream_t st) {
    auto kern = warp_variance_win_kernel<NSRC>;
    // a per-function attribute shared by every host thread and device: always the same value
    MVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    dim3 grid(cdiv(W, TW), cdiv(H, TH), B * cdiv(D, dchunk));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MVS_CUDA(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], (const uint4 *)tex16, (const float *)rt,
                                (const float *)depth_values, (uint4 *)vol_cp8, V, V - 1, p.nwin, D, H, W, dchunk, p.shp, p.win_bytes, vid));
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

}  // namespace

// tex16: fp16 RCP8 features [n_images][H][4][W][8].  view_ids_host = nullptr: n_images = B*V, image b*V + v is view v of
// batch element b; otherwise view_ids_host[b*V + v] names the image (a pool shared by several reference views).
// Asynchronous on st.
int warp_variance_windows(const void *tex16, const float *rt, const float *depth_values, void *vol_cp8, int B, int V, int D,
                          int H, int W, cudaStream_t st, int n_images, const int *view_ids_host) {
    MVS_REQUIRE(V >= 1, "warp_variance: V=%d", V);
    ViewIds vid = {};
    if (view_ids_host) {
        MVS_REQUIRE(B * V <= kMaxViewIds, "indexed features: B*V = %d exceeds %d", B * V, kMaxViewIds);
        vid.n = B * V;
        for (int i = 0; i < B * V; ++i) {
            MVS_REQUIRE(view_ids_host[i] >= 0 && view_ids_host[i] < n_images, "view id %d = %d outside the pool of %d images", i,
                        view_ids_host[i], n_images);
            vid.id[i] = view_ids_host[i];
        }
    } else {
        n_images = B * V;
    }
    int dchunk = kMaxSeg;
    MVS_REQUIRE((long long)B * cdiv(D, dchunk) <= 65535, "B*D=%lld too large for one launch", (long long)B * D);
    MVS_REQUIRE(cdiv(H, TH) <= 65535, "feature map too tall");
    const int nsrc = V - 1;
    // window plan + tensor maps depend on (features address, image count, H, W, views): cached per host thread, the
    // feature tensor of consecutive depth maps is the same workspace / pool
    struct MapSlot { const void *tex; int n, H, W, nsrc; WinPlan p; CUtensorMap maps[kShapes]; };
    static thread_local MapSlot cache[4];
    static thread_local int cache_next = 0;
    MapSlot *slot = nullptr;
    for (auto &c : cache)
        if (c.tex == tex16 && c.n == n_images && c.H == H && c.W == W && c.nsrc == nsrc) { slot = &c; break; }
    if (slot == nullptr) {
        MapSlot &c = cache[cache_next];
        cache_next = (cache_next + 1) % 4;
        c.tex = nullptr;
        c.p = plan_windows(nsrc, dchunk);
        for (int s = 0; s < kShapes; ++s)
            if (int rc = encode_window_map(&c.maps[s], tex16, n_images, H, W, c.p.shp.wx[s], c.p.shp.wy[s])) return rc;
        c.tex = tex16; c.n = n_images; c.H = H; c.W = W; c.nsrc = nsrc;
        slot = &c;
    }
    const WinPlan &p = slot->p;
    const CUtensorMap *maps = slot->maps;
    switch (p.nwin > 0 ? nsrc : 0) {
        case 1: return launch_win<1>(maps, p, tex16, rt, depth_values, vol_cp8, B, V, D, H, W, dchunk, vid, st);
        case 2: return launch_win<2>(maps, p, tex16, rt, depth_values, vol_cp8, B, V, D, H, W, dchunk, vid, st);
        case 3: return launch_win<3>(maps, p, tex16, rt, depth_values, vol_cp8, B, V, D, H, W, dchunk, vid, st);
        case 4: return launch_win<4>(maps, p, tex16, rt, depth_values, vol_cp8, B, V, D, H, W, dchunk, vid, st);
        default: return launch_win<0>(maps, p, tex16, rt, depth_values, vol_cp8, B, V, D, H, W, dchunk, vid, st);
    }
}

namespace {

constexpr int kSmemBudget32 = 225 * 1024;  // one CTA per SM

template <int NSRC>
int launch_win32(const CUtensorMap *maps, const WinPlan &p, const void *tex32, const float *rt, const float *depth_values,
                 float *var, int B, int V, int D, int H, int W, int dchunk, cudaStream_t st) {
    auto kern = warp_variance_win32_kernel<NSRC>;
    MVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget32));
    dim3 grid(cdiv(W, TW), cdiv(H, TH), B * cdiv(D, dchunk));
    kern<<<grid, kThreads32, p.smem, st>>>(maps[0], maps[1], maps[2], (const float4 *)tex32, rt, depth_values, var, V, V - 1, p.nwin,
                                           D, H, W, dchunk, p.shp, p.win_bytes);
    MVS_LAUNCH_CHECK(1);
    return MVS_OK;
}

}  // namespace

// Strict-fp32 fused warp + variance (mvs_warp_variance_fwd): fea fp32 [B,V,32,H,W] -> var fp32 [B,32,D,H,W].
// tex32: scratch for the RCP4 copy of all B*V feature maps (B*V*H*W*128 bytes).  Asynchronous on st.
int warp_variance_windows_f32(const float *fea, void *tex32, const float *rt, const float *depth_values, float *var, int B, int V,
                              int D, int H, int W, cudaStream_t st) {
    MVS_REQUIRE(V >= 1, "warp_variance: V=%d", V);
    const int dchunk = kMaxSeg;
    MVS_REQUIRE((long long)B * cdiv(D, dchunk) <= 65535, "B*D=%lld too large for one launch", (long long)B * D);
    MVS_REQUIRE(cdiv(H, TH) <= 65535, "feature map too tall");
    const int nsrc = V - 1, N = B * V;
    const long long total = (long long)N * H * 8 * W;
    nchw_to_rcp4_kernel<<<cdiv(total, 256), 256, 0, st>>>(fea, (float4 *)tex32, H, W, total);
    MVS_LAUNCH_CHECK(1);
    struct MapSlot { const void *tex; int n, H, W, nsrc; WinPlan p; CUtensorMap maps[kShapes]; };
    static thread_local MapSlot cache[4];
    static thread_local int cache_next = 0;
    MapSlot *slot = nullptr;
    for (auto &c : cache)
        if (c.tex == tex32 && c.n == N && c.H == H && c.W == W && c.nsrc == nsrc) { slot = &c; break; }
    if (slot == nullptr) {
        MapSlot &c = cache[cache_next];
        cache_next = (cache_next + 1) % 4;
        c.tex = nullptr;
        c.p = plan_windows(nsrc, dchunk, 128, kSmemBudget32, 132);
        for (int s = 0; s < kShapes; ++s)
            if (int rc = encode_window_map(&c.maps[s], tex32, N, H, W, c.p.shp.wx[s], c.p.shp.wy[s], 8)) return rc;
        c.tex = tex32; c.n = N; c.H = H; c.W = W; c.nsrc = nsrc;
        slot = &c;
    }
    const WinPlan &p = slot->p;
    switch (p.nwin > 0 ? nsrc : 0) {
        case 1: return launch_win32<1>(slot->maps, p, tex32, rt, depth_values, var, B, V, D, H, W, dchunk, st);
        case 2: return launch_win32<2>(slot->maps, p, tex32, rt, depth_values, var, B, V, D, H, W, dchunk, st);
        case 3: return launch_win32<3>(slot->maps, p, tex32, rt, depth_values, var, B, V, D, H, W, dchunk, st);
        case 4: return launch_win32<4>(slot->maps, p, tex32, rt, depth_values, var, B, V, D, H, W, dchunk, st);
        default: return launch_win32<0>(slot->maps, p, tex32, rt, depth_values, var, B, V, D, H, W, dchunk, st);
    }
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
