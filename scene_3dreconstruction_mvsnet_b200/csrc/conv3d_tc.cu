// CostRegNet on the 5th-generation tensor cores: one warp-specialised implicit-GEMM kernel
// (TMA -> shared-memory plane ring -> tcgen05.mma with TMEM accumulators -> fused epilogue) that
// covers all three layer kinds of models/mvsnet.py:33-73 through a per-layer "op table":
//   * Conv3d k3 s1 p1 (+BN+ReLU)                      conv0, conv2, conv4, conv6, prob
//   * Conv3d k3 s2 p1 (+BN+ReLU)                      conv1, conv3, conv5
//   * ConvTranspose3d k3 s2 p1 op1 (+BN+ReLU, +skip)  conv7, conv9, conv11  (8 output-parity classes)
//   * Conv2d k3 s1 p1 (+BN+ReLU), fp16 operands       FeatureNet (mvsnet.py:10-30): the N images are the planes, no z
//     taps; its two 5x5 stride-2 layers run as 3x3 stride-1 layers on a space-to-depth layout that the previous
//     layer's epilogue writes directly (featurenet_tc below)
//
// Activations live in HBM as bf16 "CP8":  [B][C/8][D][H][W][8]  -- channel chunks of 8 (16 bytes)
// are the innermost unit, so (a) a TMA box {8ch, P cols, R rows, 1 plane, C/8 chunks} lands in shared
// memory as [chunk][row][col][8ch], i.e. consecutive voxels 16 bytes apart, which is exactly the
// no-swizzle K-major UMMA operand layout (8-row core matrices of 16-byte rows); (b) im2col is free:
// the A operand of filter tap (kd,kh,kw) is the same shared-memory plane addressed at a byte offset
// (kh*P + kw)*16, and a run of 128 consecutive flattened (row, col) positions is one M=128 tile;
// (c) zero padding is TMA out-of-bounds fill; (d) stride-2 layers load the four (y,x)-parity
// sub-planes with elementStrides = 2 so that every tap is again a unit-stride operand; (e) the
// epilogue thread of TMEM lane m owns output voxel m and writes one 16-byte chunk per 8 output
// channels -- fully coalesced across the warp.
//
// A CTA (persistent, one per SM) walks a column of the volume along z: each new input plane is
// loaded exactly once into a ring of NSLOT planes and reused by the 3 (or 2) z-steps that need it.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2-9 = epilogue.
#include <algorithm>
#include <map>
#include <mutex>
#include <shared_mutex>
#include <stddef.h>
#include <string.h>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace mvs {

tmap_encode_fn get_tmap_encode() {
    static tmap_encode_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (tmap_encode_fn)p;
    });
    return fn;
}

constexpr int kMaxOps = 112;
constexpr int kMaxAcc = 8;
constexpr int kTcThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant)
constexpr int kTcThreadsDual = 352;  // + warp 10: second MMA issuer
// depth-folded kernel: warp 0 producer, warp 1 issuer, NSETS epilogue sets of 4 warps (one per TMEM lane quadrant),
// then the second issuer warp.  Measured (in the step): conv0 496 us with 2 sets, 437 us with 3; prob 210 / 160 us.
// 3 sets cap the kernel at 128 registers: fine for the kw-folded variants (119 / 128), not for the plain one (161:
// spills in the drain loop cost more than the third set gives), which keeps 2 sets.
constexpr int fold_threads(int nsets) { return 32 * (3 + 4 * nsets); }

struct TcOp {
    uint32_t a_off;     // byte offset of the A operand inside its ring plane (sub-plane + tap + chunk pair)
    uint32_t lbo;       // byte distance between the two 16-byte K chunks of this K=16 instruction
    uint16_t widx;      // index of the packed B block
    uint8_t plane_rel;  // which of the NEED planes of this step
    uint8_t acc;        // accumulator group
    // UMMA descriptor low words minus the run-time bases (filled by finalize_ops): the issue loop reads them
    // with uniform constant-bank loads, so no R2UR / shared-memory round trip sits between two tcgen05.mma
    uint32_t a_lo;      // (a_off >> 4) | (lbo >> 4) << 16
    uint32_t b_lo;      // (widx * NPAD * 32 >> 4) | (NPAD * 16 >> 4) << 16
};

struct TcLayer {
    // tile space (conv: output voxels; convT: input = low-resolution voxels)
    int B, Dt, Ht, Wt;
    int tiles_x, tiles_y, zsegs, zseg_len, ngroups, n_items;
    int TXB, TY, P, MT;
    // ring
    int nsub, chunks, sub_bytes, sub_stride, slot_bytes, need, adv, nslot, pz0, zscale;
    int in_scale, sub_xoff[4], sub_yoff[4];
    int merged_x;  // 1: tensor map is 4-D with the 8 channels and x merged into one contiguous inner dimension
    int in_split;  // 1 (stride-2 conv): the input is the (y,x)-parity-split copy [B][4][C/8][D][H/2][W/2][8] written by the
                   //    previous layer (out2): each parity sub-plane is a unit-stride box -> fast merged-x TMA
    void *out2;    // optional second output: the (y,x)-parity-split copy of `out` for a following stride-2 layer
    // gemm
    int nops, nacc, npad, wbytes_group;
    int acc_first[kMaxAcc + 1];  // op range of each accumulator group
    // issue segments: maximal runs of ops with the same accumulator group and the same ring plane
    int nseg;
    uint16_t seg_first[2 * kMaxAcc + 4], seg_last[2 * kMaxAcc + 4];
    uint8_t seg_plane[2 * kMaxAcc + 4], seg_acc[2 * kMaxAcc + 4], seg_new_acc[2 * kMaxAcc + 4];
    // epilogue
    int out_scale, acc_pz[kMaxAcc], acc_py[kMaxAcc], acc_px[kMaxAcc];
    int acc_voff[kMaxAcc];  // (pz*Hout + py)*Wout + px
    int cout_group, cout_total, relu, out_f32;
    int Dout, Hout, Wout;
    void *out;
    const uint4 *skip;
    const float *shift;
    const uint4 *wpacked;
    long long *dbg;  // optional [gridDim.x][12] cycle counters (tools/tc_profile.py); nullptr in production
    int f16;         // 1: fp16 operands / fp16 outputs (FeatureNet); 0: bf16
    int out_mode;    // 0: CP8 [C/8][D][H][W][8]; 1: space-to-depth [4 parities x C/8][D][H/2][W/2][8];
                     // 2: row-chunk-planar "RCP8" [D][H][C/8][W][8] (what the fused warp kernel's TMA windows read)
    int merged_t;    // 1: transposed conv with the 8 output-parity classes merged along N (column block = class)
    int dual;        // 1: two MMA issuer warps alternate over the steps
    int fold;        // 1: depth-folded variant (conv3d_tc_fold_kernel): the three kd taps are folded into N
    int fold_R;      // accumulator blocks per M-tile in TMEM (ring along z)
    int fold_sets;   // depth-folded kernel: number of epilogue sets (2 or 3)
    unsigned long long rcp_zsegs, rcp_tiles_x, rcp_tiles_y;  // ceil(2^40 / d) for the item decode
    // class-merged transposed conv: the skip tile of a step ([CPC][2 planes][2*TY rows][2*TXB voxels][8]) is brought to
    // shared memory by TMA one step ahead (two buffers, one per epilogue warp set) instead of being loaded by the
    // epilogue threads when they need it
    int skip_tma, skip_buf_bytes, skip_tx_bytes, skip_off;  // skip_off: offset of buffer 0 from the end of the plane ring
    alignas(64) CUtensorMap skip_map;
    int tmem_bufs_log2;  // 1 or 2: two or four accumulator buffers in TMEM
    int w_early;     // 1: the packed weights were written long before this launch (cache hit): their copy to shared memory
                     //    may start before the grid dependency wait
    int mma_n;       // > 0: N of the tcgen05.mma (kw-folded 2-D layers: 3*Cout rounded up to 16) -- the TMEM column stride
                     //      per M-tile stays the template's NPAD
    int fold_kw;     // 1: (Cin = 8, Cout = 1: the prob layer) the kw taps are folded into N as well; the epilogue adds
                     //    the three partial sums of x, x+1, x+2 (lane shifts)
    TcOp ops[kMaxOps];
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {  // saturating (finite) conversion: one F2FP instead of 4 FMNMX + F2FP
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// pack / unpack in the 16-bit storage format of the tensor-core path (kActF16, tc_common.cuh)
__device__ __forceinline__ uint32_t pack_act(float a, float b) { return kActF16 ? pack_f16x2(a, b) : pack_bf16x2(a, b); }
__device__ __forceinline__ float2 unpack_act(uint32_t u) {
    if constexpr (kActF16) return __half22float2(*reinterpret_cast<const __half2 *>(&u));
    else return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u));
}
template <bool F16>
__device__ __forceinline__ uint32_t pack16x2(float a, float b) {
    return (F16 || kActF16) ? pack_f16x2(a, b) : pack_bf16x2(a, b);
}
// ReLU on a packed pair after rounding: max(round(x), 0) == round(max(x, 0)) (rounding is monotonic, 0 is exact) -- one
// instruction for two channels instead of two fp32 max
template <bool F16>
__device__ __forceinline__ uint32_t relu16x2(uint32_t v) {
    if constexpr (F16 || kActF16) {
        const __half2 r = __hmax2(*reinterpret_cast<const __half2 *>(&v), __float2half2_rn(0.f));
        return *reinterpret_cast<const uint32_t *>(&r);
    } else {
        const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&v), __float2bfloat162_rn(0.f));
        return *reinterpret_cast<const uint32_t *>(&r);
    }
}

// floor(n / d) for n < 2^20 with rcp = ceil(2^40 / d)
__device__ __forceinline__ uint32_t fast_div40(uint32_t n, unsigned long long rcp) {
    return (uint32_t)(((unsigned long long)n * rcp) >> 40);
}

// One elected lane issues every tcgen05.mma of one z-step.  Per op: one 8-byte shared-memory load of the two
// precomputed descriptor low words, one add for the ring-slot base, MT back-to-back MMAs.
template <int NPAD, int MT>
__device__ __forceinline__ void issue_step(const TcLayer &L, const uint2 *__restrict__ optab, uint32_t tmem_buf,
                                           uint32_t sb0, uint32_t sb1, uint32_t sb2, uint32_t idesc, uint64_t desc_hi) {
    for (int sg = 0; sg < L.nseg; ++sg) {
        const uint32_t pr = L.seg_plane[sg];
        const uint32_t sb = pr == 0 ? sb0 : (pr == 1 ? sb1 : sb2);
        const uint32_t d = tmem_buf + L.seg_acc[sg] * (MT * NPAD);
        const int o0 = L.seg_first[sg], o1 = L.seg_last[sg];
        uint32_t accum = L.seg_new_acc[sg] ? 0u : 1u;
        // the table entry of the next op is loaded before this op's MMAs are pushed: its shared-memory latency and the
        // move to uniform registers overlap the time the MMAs wait for queue space (it was exposed once per loop iteration:
        // 51 instead of 39 cycles per N = 16 MMA from a single issuer)
        uint2 en = optab[o0];
#pragma unroll 2
        for (int o = o0; o < o1; ++o) {
            const uint2 e = en;
            en = optab[o + 1 < o1 ? o + 1 : o];
            const uint32_t alo = e.x + sb;
            const uint64_t bd = desc_hi | (uint64_t)e.y;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                ptx::mma_bf16_ss(d + mt * NPAD, desc_hi | (uint64_t)(alo + mt * 128), bd, idesc, accum);
            accum = 1u;
        }
    }
}

// EPI: epilogue flavour, chosen on the host.  0 = general (several accumulators per M-tile, skip, fp32 output);
// 1 = lean (one accumulator per M-tile, no skip, 16-bit output); 2 = class-merged transposed conv (+skip).
template <int NPAD, bool F16 = false, int EPI = 0>
__global__ void __launch_bounds__(kTcThreadsDual, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcLayer L) {
    constexpr bool SIMPLE = (EPI == 1);
    constexpr bool STEPWISE = (EPI != 0);  // epilogue warp sets alternate over steps (one TMEM buffer each)
    // Programmatic dependent launch: the next layer's CTAs may start on SMs this grid has left; everything up to
    // pdl_wait() below (barriers, TMEM, descriptor table) touches nothing another kernel writes.
    ptx::pdl_launch_dependents();
    extern __shared__ __align__(1024) uint8_t smem[];
    // [0,256): mbarriers + tmem address; then packed weights; then the plane ring (128-byte aligned)
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    const uint32_t bar_base = ptx::smem_u32(smem);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };   // [4] accumulator buffer b is complete
    auto tempty_bar = [&](int b) { return bar_base + 8u * (20 + b); };  // [4] ... has been drained
    auto sfull_bar = [&](int b) { return bar_base + 8u * (24 + b); };   // [2] skip tile of buffer b has landed
    auto sempty_bar = [&](int b) { return bar_base + 8u * (26 + b); };  // [2] ... has been read by the 4 warps of set b
    const uint32_t w_bar = bar_base + 8u * 28;                           // packed weights have landed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 8 * 29);
    uint2 *optab = reinterpret_cast<uint2 *>(smem + 256);  // [kMaxOps] {A desc lo (no slot base), B desc lo}
    float *s_shift = reinterpret_cast<float *>(smem + 256 + kMaxOps * 8);  // [64] folded shifts of this channel group
    constexpr uint32_t kHdr = 256 + kMaxOps * 8 + 256;
    uint8_t *w_smem = smem + kHdr;
    const uint32_t w_base = bar_base + kHdr;
    // descriptor high word: SBO = 128 B (8 rows x 16 B), version 1 (Blackwell), no swizzle
    constexpr uint64_t kDescHi = ((uint64_t)((128u >> 4) | (1u << 14))) << 32;
    const uint32_t ring_base = (w_base + L.wbytes_group + 127u) & ~127u;
    const uint32_t skip_base = (ring_base + L.nslot * L.slot_bytes + (uint32_t)L.skip_off + 1023u) & ~1023u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncols_buf = L.nacc * L.MT * NPAD;
    // 2 or 4 accumulator buffers (4 when they fit the 512 columns): steps alternate between the two issuer warps and between
    // the two epilogue warp sets by parity, so with 4 buffers each of them has two in flight and the hand-offs
    // (commit -> drain -> release -> next MMAs) of one buffer overlap the work on the other
    const uint32_t nbs = L.tmem_bufs_log2;  // log2(buffers)
    const uint32_t nbm = (1u << nbs) - 1u;
    uint32_t tmem_cols = 32;
    while (tmem_cols < ((uint32_t)ncols_buf << nbs)) tmem_cols <<= 1;

    // ---- one-time setup
    // Items are ordered so that a CTA keeps the same n-group for all of its items: group = blockIdx.x % ngroups.
    const int group = blockIdx.x % L.ngroups;
    // The packed weights (up to 110 KB) come by one bulk copy that completes on w_bar; only the MMA issuers wait for it, so
    // it overlaps the rest of the set-up and the first plane loads.  (A per-thread copy loop here was a chain of L2
    // latencies: ~2-7 us per launch depending on how far the compiler happened to unroll it.)
    (void)w_smem;
    for (int o = threadIdx.x; o < L.nops; o += blockDim.x)
        optab[o] = make_uint2(L.ops[o].a_lo, L.ops[o].b_lo + (w_base >> 4));
    if (threadIdx.x < 64)
        s_shift[threadIdx.x] = (threadIdx.x < L.cout_group) ? __ldg(L.shift + group * L.cout_group + threadIdx.x) : 0.f;
    if (threadIdx.x == 0) {
        ptx::mbar_init(w_bar, 1);
        for (int s = 0; s < L.nslot; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            // dual issuers: a plane is free when BOTH issuer warps' MMAs that read it have completed
            ptx::mbar_init(empty_bar(s), L.dual ? 2 : 1);
        }
        for (int b = 0; b < 4; ++b) {
            ptx::mbar_init(tfull_bar(b), 1);
            // lean epilogue: the 4 quadrant warps of the set that owns this buffer; general: all 8 epilogue warps
            ptx::mbar_init(tempty_bar(b), STEPWISE ? 4 : 8);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(sfull_bar(b), 1);
            ptx::mbar_init(sempty_bar(b), 4);
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&tmap);
        if (L.skip_tma) ptx::prefetch_tensormap(&L.skip_map);
    }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), tmem_cols);
    ptx::fence_proxy_async_smem();  // weights were written with generic stores, read by the MMA (async proxy)
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    // warp-uniform copy of the TMEM base (shuffle from lane 0 lets the compiler keep it in a uniform register)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    if (threadIdx.x == 0 && L.w_early) {
        ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)L.wbytes_group);
        ptx::bulk_copy_g2s(w_base, L.wpacked + (size_t)group * (L.wbytes_group / 16), (uint32_t)L.wbytes_group, w_bar);
    }
    ptx::pdl_wait();  // from here on: the previous kernel's outputs (activations, freshly packed weights) are read
    if (threadIdx.x == 0 && !L.w_early) {
        ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)L.wbytes_group);
        ptx::bulk_copy_g2s(w_base, L.wpacked + (size_t)group * (L.wbytes_group / 16), (uint32_t)L.wbytes_group, w_bar);
    }

    const int items_per_group = L.n_items / L.ngroups;
    const int cta_in_group = blockIdx.x / L.ngroups;
    const int ctas_per_group = (gridDim.x - group + L.ngroups - 1) / L.ngroups;

    auto decode = [&](int it, int &b, int &x0, int &y0, int &zs, int &T) {
        // divisions by the per-layer constants through 2^40 reciprocals (exact for it < 2^20): ~4 instructions each; an
        // item of a 2-D layer is one to five steps, and six integer divisions per item were a visible share of them
        uint32_t r = (uint32_t)it;
        uint32_t qd = fast_div40(r, L.rcp_zsegs);
        const int zseg = (int)(r - qd * (uint32_t)L.zsegs); r = qd;
        qd = fast_div40(r, L.rcp_tiles_x);
        const int tx = (int)(r - qd * (uint32_t)L.tiles_x); r = qd;
        qd = fast_div40(r, L.rcp_tiles_y);
        const int ty = (int)(r - qd * (uint32_t)L.tiles_y); r = qd;
        b = (int)r;
        x0 = tx * L.TXB;
        y0 = ty * L.TY;
        zs = zseg * L.zseg_len;
        T = min(L.zseg_len, L.Dt - zs);
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t g = 0;
            long long prod_wait = 0;
            const uint32_t tx_bytes = (uint32_t)L.nsub * L.sub_bytes;
            for (int it = cta_in_group; it < items_per_group; it += ctas_per_group) {
                int b, x0, y0, zs, T;
                decode(it, b, x0, y0, zs, T);
                const int nplanes = L.adv * (T - 1) + L.need;
                for (int j = 0; j < nplanes; ++j, ++g) {
                    const int slot = g % L.nslot;
                    const long long c0 = clock64();
                    ptx::mbar_wait(empty_bar(slot), ((g / L.nslot) & 1) ^ 1);
                    prod_wait += clock64() - c0;
                    ptx::mbar_arrive_expect_tx(full_bar(slot), tx_bytes);
                    const int pz = L.zscale * zs + L.pz0 + j;
                    if (L.merged_x) {  // inner dimension = 16*P contiguous bytes per row (uint64 elements)
                        ptx::tma_load_4d(ring_base + slot * L.slot_bytes, &tmap, full_bar(slot), 2 * (x0 + L.sub_xoff[0]),
                                         y0 + L.sub_yoff[0], pz, b * L.chunks);
                    } else if (L.in_split) {
                        for (int s = 0; s < 4; ++s)  // s = ypar*2 + xpar: its own contiguous sub-volume
                            ptx::tma_load_4d(ring_base + slot * L.slot_bytes + s * L.sub_stride, &tmap, full_bar(slot),
                                             2 * (x0 + L.sub_xoff[s]), y0 + L.sub_yoff[s], pz, (b * 4 + s) * L.chunks);
                    } else {
                        for (int s = 0; s < L.nsub; ++s)
                            ptx::tma_load_5d(ring_base + slot * L.slot_bytes + s * L.sub_stride, &tmap, full_bar(slot), 0,
                                             L.in_scale * x0 + L.sub_xoff[s], L.in_scale * y0 + L.sub_yoff[s], pz,
                                             b * L.chunks);
                    }
                }
            }
            if (L.dbg) L.dbg[blockIdx.x * 12 + 0] = prod_wait;
        } else if (lane == 1 && L.skip_tma) {
            // second producer thread: the skip tile of every step, into the buffer of the epilogue set that drains it
            uint32_t sst = 0;
            const int cpc = L.cout_total >> 3;
            for (int it = cta_in_group; it < items_per_group; it += ctas_per_group) {
                int b, x0, y0, zs, T;
                decode(it, b, x0, y0, zs, T);
                for (int t = 0; t < T; ++t, ++sst) {
                    const uint32_t sb = sst & 1u;
                    ptx::mbar_wait(sempty_bar(sb), ((sst >> 1) & 1u) ^ 1u);
                    ptx::mbar_arrive_expect_tx(sfull_bar(sb), (uint32_t)L.skip_tx_bytes);
                    ptx::tma_load_4d(skip_base + sb * L.skip_buf_bytes, &L.skip_map, sfull_bar(sb), 4 * x0, 2 * y0, 2 * (zs + t), b * cpc);
                }
            }
        }
    } else if (warp == 1 || warp == 10) {
        // ================= MMA issuer(s) =================
        // The whole warp runs this loop with warp-uniform values (kernel parameters, loop counters) so that
        // descriptors live in uniform registers; one elected lane issues tcgen05.mma / tcgen05.commit.
        // Dual mode (L.dual, warp 10 present): two issuer warps alternate over the steps -- warp 1 owns TMEM buffer 0,
        // warp 10 buffer 1.  Input planes shared by consecutive steps are released by both (empty barrier count 2).  The barrier waits, commits and
        // bookkeeping of one step (~1000 cycles, as long as its 20-70 MMAs take to execute) then overlap the other
        // warp's MMAs instead of leaving the tensor pipe idle.
        const uint32_t me = (warp == 10) ? 1u : 0u;
        const bool leader = ptx::elect_one();
        const uint32_t mma_n = L.mma_n > 0 ? (uint32_t)L.mma_n : (uint32_t)NPAD;
        const uint32_t idesc = (F16 || kActF16) ? ptx::make_idesc_f16_m128(mma_n) : ptx::make_idesc_bf16_m128(mma_n);
        uint32_t st = 0;
        uint32_t s0 = 0;    // ring slot of the oldest plane of the current step
        uint32_t par = 0;   // bit s: parity of the fill of slot s that is current (toggles when the slot is released)
        long long w_full = 0, w_tempty = 0, t_issue = 0, t_release = 0;
        const long long t_start = clock64();
        const uint32_t nslot = L.nslot, need = L.need, adv = L.adv;
        auto wrap = [&](uint32_t s) { return s >= nslot ? s - nslot : s; };
        if (leader) ptx::mbar_wait(w_bar, 0);
        // one lane runs the whole loop: no warp-level re-convergence points between a step's waits and its MMAs
        if (leader)
        for (int it = cta_in_group; it < items_per_group; it += ctas_per_group) {
            int b, x0, y0, zs, T;
            decode(it, b, x0, y0, zs, T);
            for (int t = 0; t < T; ++t, ++st) {
                const uint32_t buf = st & nbm;
                const bool mine = !L.dual || (st & 1u) == me;  // dual mode: the other warp's steps only advance the ring state
                const uint32_t sl0 = s0, sl1 = wrap(s0 + 1), sl2 = wrap(s0 + 2);
                if (leader && mine) {  // only one lane spins; the warp re-converges below so the issue loop stays uniform
                    const long long c0 = clock64();
                    // planes already waited for in earlier steps of this item need no second look (single issuer; with
                    // two issuers the previous step's waits were the other warp's)
                    for (uint32_t r = (t == 0 || L.dual) ? 0 : need - adv; r < need; ++r) {
                        const uint32_t sl = wrap(s0 + r);
                        ptx::mbar_wait(full_bar(sl), (par >> sl) & 1u);
                    }
                    const long long c1 = clock64();
                    ptx::mbar_wait(tempty_bar(buf), ((st >> nbs) & 1) ^ 1);
                    w_full += c1 - c0;
                    w_tempty += clock64() - c1;
                }
                ptx::tcgen05_fence_after();
                if (leader && mine) {
                    const long long ci = clock64();
                    const uint32_t sb0 = (ring_base + sl0 * L.slot_bytes) >> 4;
                    const uint32_t sb1 = (ring_base + sl1 * L.slot_bytes) >> 4;
                    const uint32_t sb2 = (ring_base + sl2 * L.slot_bytes) >> 4;
                    const uint32_t tbuf = tmem_base + buf * ncols_buf;
                    if (L.MT == 4) issue_step<NPAD, 4>(L, optab, tbuf, sb0, sb1, sb2, idesc, kDescHi);
                    else if (L.MT == 2) issue_step<NPAD, 2>(L, optab, tbuf, sb0, sb1, sb2, idesc, kDescHi);
                    else if (L.MT == 1) issue_step<NPAD, 1>(L, optab, tbuf, sb0, sb1, sb2, idesc, kDescHi);
                    else issue_step<NPAD, 3>(L, optab, tbuf, sb0, sb1, sb2, idesc, kDescHi);
                    t_issue += clock64() - ci;
                }
                // release the planes this step was the last user of (all of them at the end of an item)
                const long long cr = clock64();
                const uint32_t nrel = (t == T - 1) ? need : adv;
                for (uint32_t r = 0; r < nrel; ++r) {
                    const uint32_t sl = wrap(s0 + r);
                    // dual mode: both issuers arrive -- the owner of this step for its MMAs, the other one for its MMAs of
                    // the previous steps that read the plane (a commit covers everything the thread issued so far)
                    if (leader && (mine || L.dual)) ptx::tcgen05_commit(empty_bar(sl));
                    par ^= 1u << sl;
                }
                if (leader && mine) ptx::tcgen05_commit(tfull_bar(buf));
                s0 = wrap(s0 + nrel);
                t_release += clock64() - cr;
            }
        }
        if (leader && L.dbg && me == 0) {
            L.dbg[blockIdx.x * 12 + 1] = w_full;
            L.dbg[blockIdx.x * 12 + 2] = w_tempty;
            L.dbg[blockIdx.x * 12 + 3] = t_issue;
            L.dbg[blockIdx.x * 12 + 4] = clock64() - t_start;
            L.dbg[blockIdx.x * 12 + 5] = st;
            L.dbg[blockIdx.x * 12 + 8] = t_release;
        }
    } else {
        // ================= epilogue (4 warps = 128 TMEM lanes) =================
        // Per item the (row, col) of each of this thread's <= 4 M-tile rows and its output voxel offset are fixed;
        // only z advances.  Per z-step the (M-tile, accumulator) pairs are drained in batches of G = 64/NPAD:
        // all tcgen05.ld of a batch are issued back to back, the skip-connection loads of the batch are issued
        // while they fly, then one tcgen05.wait::ld -- so TMEM and global latencies overlap instead of adding up.
        constexpr int G = (NPAD <= 64) ? 64 / NPAD : 1;
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const int eset = (warp - 2) >> 2;  // the two warps of a quadrant alternate over the batches of a step
        const uint4 *skip = L.skip;
        // space-to-depth output (out_s2d): voxel (z, y, x), chunk c -> chunk ((y&1)*2 + (x&1)) * C/8 + c of a volume
        // with half the rows and columns; the parity term is folded into the per-thread base offset below
        // and RCP8 (out_mode 2) only change the chunk stride ("plane"), the z stride and the per-thread base
        const bool s2d = L.out_mode == 1, rcp8 = L.out_mode == 2;
        const size_t plane = s2d ? (size_t)L.Dout * (L.Hout / 2) * (L.Wout / 2)
                                 : (rcp8 ? (size_t)L.Wout : (size_t)L.Dout * L.Hout * L.Wout);
        const size_t zstride = s2d ? (size_t)(L.Hout / 2) * (L.Wout / 2)
                                   : (rcp8 ? (size_t)L.Hout * (L.cout_total >> 3) * L.Wout : (size_t)L.out_scale * L.Hout * L.Wout);
        const int nacc_shift = (L.nacc == 8) ? 3 : 0;
        const size_t mvoff[4] = {(size_t)L.acc_voff[0], (size_t)L.acc_voff[2], (size_t)L.acc_voff[4], (size_t)L.acc_voff[6]};
        const int npairs = L.MT << nacc_shift;
        const int nchunk = (L.cout_group + 7) >> 3;
        const int chunk0 = (group * L.cout_group) >> 3;  // first output chunk of this CTA's channel group
        // Lean path for the common layer shape (one accumulator per M-tile, no skip connection, 16-bit output): shifts in
        // registers, one output pointer per item, all tcgen05.ld of a batch of M-tiles before one wait, and the TMEM
        // buffer released as soon as its values are in registers.  ~30 instructions per (M-tile, 8 channels) instead
        // of ~290 in the general path below (transposed convs with skip, fp32 single-channel output).
        // (SIMPLE is chosen on the host: L.nacc == 1, no skip, not the fp32 single-channel output.)
        constexpr int kShr = (SIMPLE && NPAD <= 32) ? NPAD : 1;
        float shr[kShr];
#pragma unroll
        for (int e = 0; e < kShr; ++e) shr[e] = s_shift[e];
        const float relu_lo = L.relu ? 0.f : -3.0e38f;
        uint32_t st = 0;
        long long epi_wait = 0, epi_work = 0;
        // (row, col) of this thread's position in each M-tile: the same for every item (the division is hoisted out of the
        // item loop -- 2-D layers have items of one to five steps)
        uint32_t yx0 = 0, yx1 = 0, yx2 = 0, yx3 = 0;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int pos = mt * 128 + q * 32 + lane;
            const int y = pos / L.P, x = pos - y * L.P;
            const uint32_t yx = ((uint32_t)y << 16) | (uint32_t)x;
            if (mt == 0) yx0 = yx; else if (mt == 1) yx1 = yx; else if (mt == 2) yx2 = yx; else yx3 = yx;
        }
        for (int it = cta_in_group; it < items_per_group; it += ctas_per_group) {
            int b, x0, y0, zs, T;
            decode(it, b, x0, y0, zs, T);
            if (STEPWISE && T == 1 && (st & 1u) != (uint32_t)eset) {  // a one-step item that the other warp set drains
                ++st;
                continue;
            }
            size_t base0 = 0, base1 = 0, base2 = 0, base3 = 0;
            uint32_t sp0 = 0, sp1 = 0, sp2 = 0, sp3 = 0;  // per-batch offsets into the parity-split second output (L.out2)
            const size_t plane_sp = (size_t)L.Dout * (L.Hout / 2) * (L.Wout / 2);
            uint32_t vmask = 0;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                if (mt < L.MT) {
                    const uint32_t yx = mt == 0 ? yx0 : (mt == 1 ? yx1 : (mt == 2 ? yx2 : yx3));
                    const int y = (int)(yx >> 16), x = (int)(yx & 0xffffu);
                    const bool valid = (y < L.TY) && (x < L.TXB) && (y0 + y < L.Ht) && (x0 + x < L.Wt);
                    if (L.skip_tma) {  // byte offset of voxel (2y, 2x) inside a plane of the skip tile in shared memory
                        const uint32_t sp = ((uint32_t)(2 * y) * 2u * (uint32_t)L.TXB + (uint32_t)(2 * x)) * 16u;
                        if (mt == 0) sp0 = sp; else if (mt == 1) sp1 = sp; else if (mt == 2) sp2 = sp; else sp3 = sp;
                    } else if (L.out2 != nullptr) {
                        const uint32_t sp = (uint32_t)((((y0 + y) & 1) * 2 + ((x0 + x) & 1)) * (L.cout_total >> 3)) * (uint32_t)plane_sp +
                                            (uint32_t)((y0 + y) >> 1) * (L.Wout / 2) + (uint32_t)((x0 + x) >> 1);
                        if (mt == 0) sp0 = sp; else if (mt == 1) sp1 = sp; else if (mt == 2) sp2 = sp; else sp3 = sp;
                    }
                    size_t bs = (size_t)(L.out_scale * (y0 + y)) * L.Wout + (size_t)L.out_scale * (x0 + x);
                    if (s2d)
                        bs = (size_t)((((y0 + y) & 1) * 2 + ((x0 + x) & 1)) * (L.cout_total >> 3)) * plane +
                             (size_t)((y0 + y) >> 1) * (L.Wout / 2) + (size_t)((x0 + x) >> 1);
                    if (rcp8) bs = (size_t)(y0 + y) * (L.cout_total >> 3) * L.Wout + (size_t)(x0 + x);
                    vmask |= (valid ? 1u : 0u) << mt;
                    if (mt == 0) base0 = bs; else if (mt == 1) base1 = bs; else if (mt == 2) base2 = bs; else base3 = bs;
                }
            }
            for (int t = 0; t < T; ++t, ++st) {
                // Lean path: the two warps of a quadrant alternate over the STEPS (epilogue set e drains TMEM buffer e),
                // so each warp has two step times for one drain and the per-step fixed costs (barrier wait,
                // tcgen05.wait::ld round trip, release) are paid once per step.  General path (skip loads in the loop):
                // both warps work on every step and split its (M-tile, accumulator) pairs -- measured faster there.
                const uint32_t buf = st & nbm;
                const uint32_t sbuf = st & 1u;  // skip-tile buffer (two of them, one per warp set)
                if (STEPWISE && sbuf != (uint32_t)eset) continue;
                const long long c0 = clock64();
                ptx::mbar_wait(tfull_bar(buf), (st >> nbs) & 1);
                const long long c1 = clock64();
                epi_wait += c1 - c0;
                ptx::tcgen05_fence_after();
                const size_t zoff = (size_t)(zs + t) * zstride;
                const uint32_t tbuf = tmem_base + ((uint32_t)(q * 32) << 16) + buf * ncols_buf;
                if constexpr (EPI == 3) {
                    // kw-folded 2-D layer (FeatureNet): column kw*COUT + co of an M-tile is
                    // U_kw[p][co] = sum_{kh,ci} in[p + kh*P][ci] w[co][ci][kh][kw]; out[p] = U_0[p] + U_1[p+1] + U_2[p+2].
                    // The row pitch is P = 32 positions = one 32-lane group, and the last two positions of a row are halo
                    // (never an output), so the shift never leaves the warp: two shuffles per channel, no edge exchange.
                    constexpr int COUT = (NPAD == 32) ? 8 : (NPAD == 64 ? 16 : 32);
                    constexpr int MB = (COUT == 8) ? 2 : 1;  // M-tiles per batch
                    uint4 *oz = reinterpret_cast<uint4 *>(L.out) + zoff;
                    for (int m0 = 0; m0 < L.MT; m0 += MB) {
                        uint32_t r[MB][3 * COUT];
#pragma unroll
                        for (int i = 0; i < MB; ++i)
                            if (m0 + i < L.MT) {
#pragma unroll
                                for (int c8 = 0; c8 < 3 * COUT / 8; ++c8) ptx::tmem_ld_x8(tbuf + (m0 + i) * NPAD + c8 * 8, &r[i][c8 * 8]);
                            }
                        ptx::tmem_ld_wait();
                        if (m0 + MB >= L.MT) {  // everything this warp needs from the buffer is in registers
                            ptx::tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) ptx::mbar_arrive(tempty_bar(buf));
                        }
#pragma unroll
                        for (int i = 0; i < MB; ++i) {
                            const int mt = m0 + i;
                            if (mt >= L.MT) continue;  // warp-uniform
                            uint32_t pk[COUT / 2];
#pragma unroll
                            for (int c = 0; c < COUT; c += 2) {  // two channels at a time: packed fp32 adds, same order as before
                                const float2 a = make_float2(__uint_as_float(r[i][c]), __uint_as_float(r[i][c + 1]));
                                const float2 u1 = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(r[i][COUT + c]), 1),
                                                              __shfl_down_sync(0xffffffffu, __uint_as_float(r[i][COUT + c + 1]), 1));
                                const float2 u2 = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(r[i][2 * COUT + c]), 2),
                                                              __shfl_down_sync(0xffffffffu, __uint_as_float(r[i][2 * COUT + c + 1]), 2));
                                const float2 v = __fadd2_rn(__fadd2_rn(__fadd2_rn(a, u1), u2), make_float2(s_shift[c], s_shift[c + 1]));
                                const uint32_t p = pack16x2<F16>(v.x, v.y);
                                pk[c / 2] = L.relu ? relu16x2<F16>(p) : p;
                            }
                            if (!((vmask >> mt) & 1u)) continue;
                            uint4 *op = oz + (mt == 0 ? base0 : (mt == 1 ? base1 : (mt == 2 ? base2 : base3)));
#pragma unroll
                            for (int c8 = 0; c8 < COUT / 8; ++c8)
                                op[(size_t)c8 * plane] = make_uint4(pk[c8 * 4], pk[c8 * 4 + 1], pk[c8 * 4 + 2], pk[c8 * 4 + 3]);
                        }
                    }
                    epi_work += clock64() - c1;
                } else if constexpr (EPI == 2) {
                    // Class-merged transposed conv: the M-tile's NPAD columns are [class (pz,py,px)][Cout].  A "pair" is
                    // the two x-parity classes of one (pz, py, channel chunk): their output chunks are adjacent in
                    // memory (voxels 2x and 2x+1), so each thread moves 32 contiguous bytes with one 256-bit load /
                    // store and a warp covers 1 KB per instruction.  Everything is unrolled over compile-time
                    // (pz, py, chunk); the skip loads of a group are issued before its tcgen05.ld so both latencies
                    // overlap.  (A bulk L2 prefetch of the skip rows by the producer warp was measured: 6 % L2 hit rate and
                    // +330 MB of DRAM reads -- the lines are gone before they are used -- so it was removed.)
                    constexpr int COUT = NPAD / 8, CPC = COUT / 8;
                    const uint4 *skp = skip ? skip + zoff : nullptr;
                    uint4 *outp = reinterpret_cast<uint4 *>(L.out) + zoff;
                    // skip tile in shared memory (L.skip_tma): voxel (pz, 2y+py, 2x+px) of chunk cc at
                    // (((cc*2 + pz) * 2TY + 2y+py) * 2TXB + 2x+px) * 16 bytes; sp* hold the thread's (2y, 2x) part
                    const uint32_t s_row = 2u * (uint32_t)L.TXB * 16u, s_plane = 2u * (uint32_t)L.TY * s_row;
                    const uint32_t s_buf = skip_base + sbuf * (uint32_t)L.skip_buf_bytes;
                    if (L.skip_tma) ptx::mbar_wait(sfull_bar(sbuf), (st >> 1) & 1);
                    for (int mt = 0; mt < L.MT; ++mt) {
                        const bool ok = (vmask >> mt) & 1u;
                        const size_t mb = mt == 0 ? base0 : (mt == 1 ? base1 : (mt == 2 ? base2 : base3));
                        const uint32_t s_thr = s_buf + (mt == 0 ? sp0 : (mt == 1 ? sp1 : (mt == 2 ? sp2 : sp3)));
#pragma unroll
                        for (int pz = 0; pz < 2; ++pz) {
                            uint32_t sk[2 * CPC][8];
                            if (skp != nullptr && ok) {
                                if (L.skip_tma) {
#pragma unroll
                                    for (int py = 0; py < 2; ++py)
#pragma unroll
                                        for (int cc = 0; cc < CPC; ++cc) {
                                            const uint32_t a = s_thr + (uint32_t)(cc * 2 + pz) * s_plane + (uint32_t)py * s_row;
                                            uint32_t *d = sk[py * CPC + cc];
                                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a));
                                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]) : "r"(a + 16u));
                                        }
                                } else {
#pragma unroll
                                    for (int py = 0; py < 2; ++py)
#pragma unroll
                                        for (int cc = 0; cc < CPC; ++cc)
                                            ptx::ldg256(skp + ((size_t)b * CPC + cc) * plane + mb + mvoff[pz * 2 + py], sk[py * CPC + cc]);
                                }
                            }
                            uint32_t r[2 * CPC][16];
#pragma unroll
                            for (int py = 0; py < 2; ++py)
#pragma unroll
                                for (int cc = 0; cc < CPC; ++cc) {
                                    const int col0 = 2 * (pz * 2 + py) * COUT + cc * 8;
                                    ptx::tmem_ld_x8(tbuf + mt * NPAD + col0, &r[py * CPC + cc][0]);
                                    ptx::tmem_ld_x8(tbuf + mt * NPAD + col0 + COUT, &r[py * CPC + cc][8]);
                                }
                            ptx::tmem_ld_wait();
                            if (mt == L.MT - 1 && pz == 1) {  // last group: the buffer can be refilled
                                ptx::tcgen05_fence_before();
                                __syncwarp();
                                if (lane == 0) ptx::mbar_arrive(tempty_bar(buf));
                            }
                            if (!ok) continue;
#pragma unroll
                            for (int py = 0; py < 2; ++py)
#pragma unroll
                                for (int cc = 0; cc < CPC; ++cc) {
                                    const int u = py * CPC + cc;
                                    uint32_t pk[8];
#pragma unroll
                                    for (int h = 0; h < 2; ++h) {  // h = x parity; two channels at a time (packed fp32 adds)
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            float2 a = __fadd2_rn(make_float2(__uint_as_float(r[u][h * 8 + 2 * e]), __uint_as_float(r[u][h * 8 + 2 * e + 1])),
                                                                  make_float2(s_shift[cc * 8 + 2 * e], s_shift[cc * 8 + 2 * e + 1]));
                                            a.x = fmaxf(a.x, relu_lo);
                                            a.y = fmaxf(a.y, relu_lo);
                                            if (skp != nullptr)
                                                a = __fadd2_rn(a, unpack_act(sk[u][h * 4 + e]));
                                            pk[h * 4 + e] = pack_act(a.x, a.y);
                                        }
                                    }
                                    ptx::stg256(outp + ((size_t)b * CPC + cc) * plane + mb + mvoff[pz * 2 + py], pk);
                                }
                        }
                    }
                    if (L.skip_tma) {  // every skip value of this step is in registers or stored: the buffer can be refilled
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(sempty_bar(sbuf));
                    }
                    epi_work += clock64() - c1;
                } else if constexpr (SIMPLE) {
                    uint4 *oz = reinterpret_cast<uint4 *>(L.out) + ((size_t)b * (L.cout_total >> 3) + chunk0) * plane + zoff;
                    constexpr int MB = (NPAD <= 16) ? 4 : (NPAD <= 32 ? 2 : 1);  // M-tiles per batch: 64 accumulator registers
                    for (int m0 = 0; m0 < L.MT; m0 += MB) {
                        uint32_t r[MB][NPAD];
#pragma unroll
                        for (int i = 0; i < MB; ++i)
                            if (m0 + i < L.MT) {
#pragma unroll
                                for (int c8 = 0; c8 < NPAD / 8; ++c8)
                                    if (c8 < nchunk) ptx::tmem_ld_x8(tbuf + (m0 + i) * NPAD + c8 * 8, &r[i][c8 * 8]);
                            }
                        ptx::tmem_ld_wait();
                        if (m0 + MB >= L.MT) {  // everything this warp needs from the buffer is in registers
                            ptx::tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) ptx::mbar_arrive(tempty_bar(buf));
                        }
#pragma unroll
                        for (int i = 0; i < MB; ++i) {
                            const int mt = m0 + i;
                            if (mt >= L.MT || !((vmask >> mt) & 1u)) continue;
                            uint4 *op = oz + (mt == 0 ? base0 : (mt == 1 ? base1 : (mt == 2 ? base2 : base3)));
#pragma unroll
                            for (int c8 = 0; c8 < NPAD / 8; ++c8) {
                                if (c8 >= nchunk) break;
                                uint32_t pw[4];
#pragma unroll
                                for (int e = 0; e < 8; e += 2) {  // packed fp32 add of the shift, ReLU on the packed 16-bit pair
                                    const float2 sh = NPAD <= 32 ? make_float2(shr[(c8 * 8 + e) % kShr], shr[(c8 * 8 + e + 1) % kShr])
                                                                 : make_float2(s_shift[c8 * 8 + e], s_shift[c8 * 8 + e + 1]);
                                    const float2 v = __fadd2_rn(make_float2(__uint_as_float(r[i][c8 * 8 + e]), __uint_as_float(r[i][c8 * 8 + e + 1])), sh);
                                    const uint32_t p = pack16x2<F16>(v.x, v.y);
                                    pw[e / 2] = L.relu ? relu16x2<F16>(p) : p;
                                }
                                const uint4 pk = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                                op[(size_t)c8 * plane] = pk;
                                if (L.out2 != nullptr)
                                    reinterpret_cast<uint4 *>(L.out2)[((size_t)b * 4 * (L.cout_total >> 3) + chunk0 + c8) * plane_sp +
                                                                      (mt == 0 ? sp0 : (mt == 1 ? sp1 : (mt == 2 ? sp2 : sp3))) +
                                                                      (size_t)(zs + t) * ((size_t)(L.Hout / 2) * (L.Wout / 2))] = pk;
                            }
                        }
                    }
                    epi_work += clock64() - c1;
                } else if constexpr (NPAD <= 64) {
                const int gsz = min(G, (npairs + 1) >> 1);
                for (int p0 = eset * gsz; p0 < npairs; p0 += 2 * gsz) {
                    uint32_t r[G][NPAD];
                    uint4 sk[G][NPAD / 8];
                    size_t vox[G];
                    bool ok[G];
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        const int pr = p0 + j;
                        ok[j] = false;
                        vox[j] = 0;
                        if (j < gsz && pr < npairs) {  // warp-uniform
                            const int mt = pr >> nacc_shift, a = pr & (L.nacc - 1);
#pragma unroll
                            for (int c8 = 0; c8 < NPAD / 8; ++c8)
                                if (c8 < nchunk) ptx::tmem_ld_x8(tbuf + (a * L.MT + mt) * NPAD + c8 * 8, &r[j][c8 * 8]);
                            const size_t bs = mt == 0 ? base0 : (mt == 1 ? base1 : (mt == 2 ? base2 : base3));
                            ok[j] = (vmask >> mt) & 1u;
                            vox[j] = bs + zoff + (size_t)L.acc_voff[a];
                        }
                    }
                    if (skip != nullptr) {
#pragma unroll
                        for (int j = 0; j < G; ++j)
#pragma unroll
                            for (int c8 = 0; c8 < NPAD / 8; ++c8)
                                if (ok[j] && c8 < nchunk)
                                    sk[j][c8] = __ldg(skip + ((size_t)b * (L.cout_total / 8) + chunk0 + c8) * plane + vox[j]);
                    }
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        if (!ok[j]) continue;
#pragma unroll
                        for (int c8 = 0; c8 < NPAD / 8; ++c8) {
                            if (c8 >= nchunk) break;
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                v[e] = __uint_as_float(r[j][c8 * 8 + e]) + s_shift[c8 * 8 + e];
                                if (L.relu) v[e] = fmaxf(v[e], 0.f);
                            }
                            if (L.out_f32) {  // single-channel fp32 output (the prob layer): [B][D][H][W]
                                reinterpret_cast<float *>(L.out)[(size_t)b * plane + vox[j]] = v[0];
                            } else {
                                if (skip != nullptr) {  // skip + relu(bn(convT))  (mvsnet.py:69-71)
                                    const uint32_t *sp = reinterpret_cast<const uint32_t *>(&sk[j][c8]);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float2 f = unpack_act(sp[e]);
                                        v[2 * e] += f.x;
                                        v[2 * e + 1] += f.y;
                                    }
                                }
                                uint4 pk;
                                pk.x = pack16x2<F16>(v[0], v[1]);
                                pk.y = pack16x2<F16>(v[2], v[3]);
                                pk.z = pack16x2<F16>(v[4], v[5]);
                                pk.w = pack16x2<F16>(v[6], v[7]);
                                reinterpret_cast<uint4 *>(L.out)[((size_t)b * (L.cout_total / 8) + chunk0 + c8) * plane + vox[j]] = pk;
                            }
                        }
                    }
                }
                ptx::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(buf));
                epi_work += clock64() - c1;
                }  // general path
            }
        }
        if (warp == 2 && lane == 0 && L.dbg) {
            L.dbg[blockIdx.x * 12 + 6] = epi_wait;
            L.dbg[blockIdx.x * 12 + 7] = epi_work;
        }
    }

    // ---- teardown
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, tmem_cols);
    }
    (void)bars;
}

// ------------------------------------------------------------------------------------------------
// Depth-folded variant for stride-1 layers with Cout <= 16 (conv0 32->8, conv2 16->16, prob 8->1).
//
// With N = 16 every K=16 tcgen05.mma streams 4 KB of activations from shared memory for 16 output columns and
// costs ~39 cycles regardless (tools/mma_microbench2.cu: 32 + N/4 cycles up to N = 128): conv0 spent 54 of them
// per 128-voxel tile and was bound by that operand path (profiles/r01e).  Here the CTA walks INPUT planes instead
// of output planes: input plane p contributes to output planes p-1, p, p+1 through the taps kd = 2, 1, 0, so one
// MMA per (kh, kw, K-chunk) with N = 3*CW computes all three contributions at once (44 cycles at N = 48) --
// 18 MMAs per plane for conv0 instead of 54.  The accumulators of consecutive output planes are consecutive
// CW-column blocks of a ring of R blocks in TMEM (per M-tile), so the three contributions land in the right
// accumulators by construction: same TMEM lane (= voxel), adjacent column blocks.  Every instruction accumulates:
// the epilogue zeroes a block (tcgen05.st) right after draining it, so the issue loop stays as lean as the
// output-stationary one (one descriptor add per MMA, everything else loop-invariant).  Windows never wrap: the ring
// of R logical blocks lives in R+2 physical blocks (see the kernel); partial windows at the ends of a z-segment are a
// B-row offset plus a smaller N.  An output plane is complete -- and handed to the epilogue through its own mbarrier --
// one step after its centre plane.  Each input plane is used by exactly one step, so the plane ring is pure TMA
// prefetch depth.
// ------------------------------------------------------------------------------------------------
constexpr int kFoldMaxR = 16;

template <int MT>
__device__ __forceinline__ void issue_fold_pass(const uint2 *__restrict__ optab, int nops, uint32_t d, uint32_t cols_mt,
                                                uint32_t sb, uint32_t brow, uint32_t idesc, uint64_t desc_hi) {
    // (loading the table entry one op ahead as in issue_step was measured here too: conv0 442 -> 477 us -- not kept)
#pragma unroll 2
    for (int o = 0; o < nops; ++o) {
        const uint2 e = optab[o];
        const uint32_t alo = e.x + sb;
        const uint64_t bd = desc_hi | (uint64_t)(e.y + brow);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
            ptx::mma_bf16_ss(d + mt * cols_mt, desc_hi | (uint64_t)(alo + mt * 128), bd, idesc, 1u);
    }
}

template <int CW, int NSETS, bool KWF>  // KWF: kw folded into N too (CW = 16: prob, Cout = 1; CW = 32: conv0, Cout = 8)
__global__ void __launch_bounds__(fold_threads(NSETS), 1)
conv3d_tc_fold_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcLayer L) {
    ptx::pdl_launch_dependents();  // see conv3d_tc_kernel
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t bar_base = ptx::smem_u32(smem);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (32 + b); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 384);
    uint2 *optab = reinterpret_cast<uint2 *>(smem + 400);
    float *s_shift = reinterpret_cast<float *>(smem + 400 + kMaxOps * 8);  // [64]
    constexpr uint32_t kHdr = 1664;
    static_assert(400 + kMaxOps * 8 + 256 <= kHdr, "fold kernel header overflow");
    uint8_t *w_smem = smem + kHdr;
    const uint32_t w_base = bar_base + kHdr;
    constexpr uint64_t kDescHi = ((uint64_t)((128u >> 4) | (1u << 14))) << 32;
    const uint32_t ring_base = (w_base + L.wbytes_group + 127u) & ~127u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MT = L.MT;
    // Accumulator blocks per M-tile: PB physical blocks hold a ring of R = PB - 2 logical blocks.  The window of step j
    // is always the physical blocks [p, p+1, p+2], p = logical ring position of its first block, so it never wraps: when
    // p = R-2 or R-1 its tail lands in the two extra blocks R, R+1, which are "mirrors" of logical blocks 0 and 1 -- the
    // epilogue adds a mirror to its block when it drains it.  One MMA per op and M-tile in every step.
    const int PB = L.fold_R;
    const int R = PB - 2;
    const uint32_t cols_mt = PB * CW;       // TMEM columns per M-tile
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)MT * cols_mt) tmem_cols <<= 1;

    (void)w_smem;  // packed weights: one bulk copy completing on w_bar, awaited by the issuers only (see conv3d_tc_kernel)
    const uint32_t w_bar = bar_base + 392u;
    for (int o = threadIdx.x; o < L.nops; o += blockDim.x)
        optab[o] = make_uint2(L.ops[o].a_lo, L.ops[o].b_lo + (w_base >> 4));
    if (threadIdx.x < 64) s_shift[threadIdx.x] = (threadIdx.x < L.cout_group) ? __ldg(L.shift + threadIdx.x) : 0.f;
    if (threadIdx.x == 0) {
        ptx::mbar_init(w_bar, 1);
        for (int s = 0; s < L.nslot; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 2);  // both issuer warps release a plane
        }
        for (int b = 0; b < R; ++b) {
            ptx::mbar_init(tfull_bar(b), 2);  // both issuer warps complete a block (each its M-tiles)
            ptx::mbar_init(tempty_bar(b), 4);  // the 4 quadrant warps of the epilogue set that drains the block
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&tmap);
    }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), tmem_cols);
    ptx::fence_proxy_async_smem();
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // all accumulator blocks start from zero (afterwards the epilogue re-zeroes each block as it drains it)
    if (warp >= 2 && warp < 6) {
        const uint32_t tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (uint32_t c = 0; c < (uint32_t)MT * cols_mt; c += 16) ptx::tmem_st_zero_x16(tb + c);
        ptx::tmem_st_wait();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();

    if (threadIdx.x == 0 && L.w_early) {
        ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)L.wbytes_group);
        ptx::bulk_copy_g2s(w_base, L.wpacked, (uint32_t)L.wbytes_group, w_bar);
    }
    ptx::pdl_wait();  // the previous kernel's outputs are read from here on
    if (threadIdx.x == 0 && !L.w_early) {
        ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)L.wbytes_group);
        ptx::bulk_copy_g2s(w_base, L.wpacked, (uint32_t)L.wbytes_group, w_bar);
    }

    auto decode = [&](int it, int &b, int &x0, int &y0, int &zs, int &T) {
        // divisions by the per-layer constants through 2^40 reciprocals (exact for it < 2^20): ~4 instructions each; an
        // item of a 2-D layer is one to five steps, and six integer divisions per item were a visible share of them
        uint32_t r = (uint32_t)it;
        uint32_t qd = fast_div40(r, L.rcp_zsegs);
        const int zseg = (int)(r - qd * (uint32_t)L.zsegs); r = qd;
        qd = fast_div40(r, L.rcp_tiles_x);
        const int tx = (int)(r - qd * (uint32_t)L.tiles_x); r = qd;
        qd = fast_div40(r, L.rcp_tiles_y);
        const int ty = (int)(r - qd * (uint32_t)L.tiles_y); r = qd;
        b = (int)r;
        x0 = tx * L.TXB;
        y0 = ty * L.TY;
        zs = zseg * L.zseg_len;
        T = min(L.zseg_len, L.Dt - zs);
    };

    if (warp == 0) {
        // ================= TMA producer: input planes zs-1 .. zs+T of every item =================
        if (lane == 0) {
            uint32_t g = 0;
            long long prod_wait = 0;
            const uint32_t tx_bytes = (uint32_t)L.sub_bytes;
            for (int it = blockIdx.x; it < L.n_items; it += gridDim.x) {
                int b, x0, y0, zs, T;
                decode(it, b, x0, y0, zs, T);
                for (int j = 0; j < T + 2; ++j, ++g) {
                    const int slot = g % L.nslot;
                    const long long c0 = clock64();
                    ptx::mbar_wait(empty_bar(slot), ((g / L.nslot) & 1) ^ 1);
                    prod_wait += clock64() - c0;
                    ptx::mbar_arrive_expect_tx(full_bar(slot), tx_bytes);
                    ptx::tma_load_4d(ring_base + slot * L.slot_bytes, &tmap, full_bar(slot), 2 * (x0 - 1), y0 - 1, zs - 1 + j,
                                     b * L.chunks);
                }
            }
            if (L.dbg) L.dbg[blockIdx.x * 12 + 0] = prod_wait;
        }
    } else if (warp == 1 || warp == 2 + 4 * NSETS) {
        // ================= MMA issuers =================
        // Two issuer warps split the M-TILES of every step (disjoint TMEM column regions, so no accumulator is shared between
        // them): the barrier waits and commits of one overlap the other's MMAs.  Both arrive on the plane's empty barrier
        // and on the block's full barrier.
        const int mt_lo = (warp == 1) ? 0 : (MT + 1) / 2, mt_n = (warp == 1) ? (MT + 1) / 2 : MT / 2;
        const bool leader = ptx::elect_one();
        uint32_t g = 0;      // input planes consumed so far
        uint32_t slot = 0;   // ring slot of plane g, and the parity of its current fill (all tracked incrementally:
        uint32_t fpar = 0;   // no division or modulo in the per-plane path)
        uint32_t pj = 0;     // TMEM ring position of accumulator block j of the current item
        uint32_t epar = 0;   // bit b: parity of tempty_bar(b) to wait for next
        long long w_full = 0, w_tempty = 0, t_issue = 0;
        const long long t_start = clock64();
        const int nops = L.nops;
        const uint32_t nslot = L.nslot, uR = (uint32_t)R;
        auto wrapR = [&](uint32_t v) { return v >= uR ? v - uR : v; };
        if (leader) ptx::mbar_wait(w_bar, 0);
        // one lane runs the whole loop: no warp-level re-convergence points between a step's waits and its MMAs
        if (leader)
        for (int it = blockIdx.x; it < L.n_items; it += gridDim.x) {
            int b_, x0, y0, zs, T;
            decode(it, b_, x0, y0, zs, T);
            for (int j = 0; j < T + 2; ++j, ++g) {
                // B row block k (kd = 2 - k) of input plane j feeds accumulator block j + k; valid outputs are 2 .. T+1
                const int k0 = max(0, 2 - j), k1 = min(2, T + 1 - j);
                const bool first_touch = (k1 == 2);
                const uint32_t ft_blk = wrapR(pj + 2);
                if (leader) {
                    const long long c0 = clock64();
                    ptx::mbar_wait(full_bar(slot), fpar);
                    const long long c1 = clock64();
                    // the block this plane touches first must have been drained and re-zeroed by the epilogue
                    if (first_touch) ptx::mbar_wait(tempty_bar(ft_blk), ((epar >> ft_blk) & 1u) ^ 1u);
                    w_full += c1 - c0;
                    w_tempty += clock64() - c1;
                }
                if (first_touch) epar ^= 1u << ft_blk;
                ptx::tcgen05_fence_after();
                const int cnt = k1 - k0 + 1;
                if (leader) {
                    const long long ci = clock64();
                    const uint32_t sb = (ring_base + slot * L.slot_bytes) >> 4;
                    const uint32_t d = tmem_base + (pj + k0) * CW + mt_lo * cols_mt;  // window start: never wraps (mirror blocks)
                    const uint32_t brow = k0 * CW;
                    const uint32_t idesc = kActF16 ? ptx::make_idesc_f16_m128(cnt * CW) : ptx::make_idesc_bf16_m128(cnt * CW);
                    const uint32_t sbm = sb + mt_lo * 128;  // this warp's first M-tile
                    if (mt_n == 2) issue_fold_pass<2>(optab, nops, d, cols_mt, sbm, brow, idesc, kDescHi);
                    else if (mt_n == 1) issue_fold_pass<1>(optab, nops, d, cols_mt, sbm, brow, idesc, kDescHi);
                    t_issue += clock64() - ci;
                    ptx::tcgen05_commit(empty_bar(slot));
                    if (j >= 2) ptx::tcgen05_commit(tfull_bar(pj));  // output plane zs + j - 2 is complete
                }
                if (++slot == nslot) { slot = 0; fpar ^= 1u; }
                pj = wrapR(pj + 1);
            }
        }
        if (leader && L.dbg && warp == 1) {
            L.dbg[blockIdx.x * 12 + 1] = w_full;
            L.dbg[blockIdx.x * 12 + 2] = w_tempty;
            L.dbg[blockIdx.x * 12 + 3] = t_issue;
            L.dbg[blockIdx.x * 12 + 4] = clock64() - t_start;
            L.dbg[blockIdx.x * 12 + 5] = g;
        }
    } else {
        // ================= epilogue: NSETS sets of 4 warps (one per TMEM lane quadrant) =================
        // The sets take the OUTPUT PLANES in turn: set e drains every NSETS-th completed accumulator block, all M-tiles
        // of its quadrant, so the fixed costs of a drain (barrier wait, tcgen05.ld / st round trips, release) are paid once
        // per plane and quadrant and NSETS planes are in the epilogue at any time.  A drain is a long chain of dependent
        // instructions in one warp (~7 cycles each), so its throughput scales with the number of sets, not with ILP.
        const int q = warp & 3;
        const int eset = (warp - 2) >> 2;
        const size_t plane = (size_t)L.Dout * L.Hout * L.Wout;
        const size_t zstride = (size_t)L.Hout * L.Wout;
        const int nchunk = (L.cout_group + 7) >> 3;
        uint32_t blk = 2 % R, fpar = 0;  // ring position of the next block to drain (block 2 of the first item)
        uint32_t eturn = 0;              // whose turn the next output plane is
        long long epi_wait = 0, epi_work = 0;
        for (int it = blockIdx.x; it < L.n_items; it += gridDim.x) {
            int b, x0, y0, zs, T;
            decode(it, b, x0, y0, zs, T);
            size_t base[4] = {0, 0, 0, 0};
            uint32_t sp[4] = {0, 0, 0, 0};  // per-batch offsets into the parity-split second output (L.out2)
            const size_t plane_sp = (size_t)L.Dout * (L.Hout / 2) * (L.Wout / 2), zstride_sp = (size_t)(L.Hout / 2) * (L.Wout / 2);
            uint32_t vmask = 0;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                if (mt < MT) {
                    const int pos = mt * 128 + q * 32 + lane;
                    const int y = pos / L.P, x = pos - y * L.P;
                    if ((y < L.TY) && (x < L.TXB) && (y0 + y < L.Ht) && (x0 + x < L.Wt)) vmask |= 1u << mt;
                    base[mt] = (size_t)(y0 + y) * L.Wout + (size_t)(x0 + x);
                    sp[mt] = (uint32_t)((((y0 + y) & 1) * 2 + ((x0 + x) & 1)) * (L.cout_total >> 3)) * (uint32_t)plane_sp +
                             (uint32_t)((y0 + y) >> 1) * (L.Wout / 2) + (uint32_t)((x0 + x) >> 1);
                }
            }
            for (int e = 0; e < T; ++e) {
                const uint32_t myblk = blk;
                if (++blk == (uint32_t)R) blk = 0;
                const uint32_t par = (fpar >> myblk) & 1u;
                fpar ^= 1u << myblk;  // every completion of the block flips its phase, whichever set drains it
                const bool mine = (eturn == (uint32_t)eset);
                if (++eturn == (uint32_t)NSETS) eturn = 0;
                if (!mine) continue;  // another set's plane
                const long long c0 = clock64();
                ptx::mbar_wait(tfull_bar(myblk), par);
                const long long c1 = clock64();
                epi_wait += c1 - c0;
                ptx::tcgen05_fence_after();
                const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + myblk * CW;
                const bool mirrored = myblk < 2;  // logical blocks 0 and 1 have a second part in physical blocks R, R+1
                const size_t zoff = (size_t)(zs + e) * zstride;
                if constexpr (CW == 32 && KWF) {
                    // conv0 (Cout = 8, kw folded into N, MT <= 2): column kw*8 + co of a block is
                    // U_kw[p][co] = sum_{kh,ci} in[p + kh*P][ci] w[co][ci][kh][kw]; out[p] = U_0[p] + U_1[p+1] + U_2[p+2]
                    // (columns 24..31 have zero weights).  Same lane shifts as the prob layer below, eight channels wide.
                    uint32_t r[2][24];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
                        if (mt < MT) {
#pragma unroll
                            for (int c8 = 0; c8 < 3; ++c8) ptx::tmem_ld_x8(tb + mt * cols_mt + c8 * 8, &r[mt][c8 * 8]);
                        }
                    ptx::tmem_ld_wait();
                    if (mirrored) {
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
                            if (mt < MT) {
                                uint32_t r2[24];
#pragma unroll
                                for (int c8 = 0; c8 < 3; ++c8) ptx::tmem_ld_x8(tb + R * CW + mt * cols_mt + c8 * 8, &r2[c8 * 8]);
                                ptx::tmem_ld_wait();
#pragma unroll
                                for (int k = 0; k < 24; ++k) r[mt][k] = __float_as_uint(__uint_as_float(r[mt][k]) + __uint_as_float(r2[k]));
                            }
                    }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
                        if (mt < MT) {
                            ptx::tmem_st_zero_x16(tb + mt * cols_mt);
                            ptx::tmem_st_zero_x16(tb + mt * cols_mt + 16);
                            if (mirrored) {
                                ptx::tmem_st_zero_x16(tb + R * CW + mt * cols_mt);
                                ptx::tmem_st_zero_x16(tb + R * CW + mt * cols_mt + 16);
                            }
                        }
                    ptx::tmem_st_wait();
                    ptx::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(tempty_bar(myblk));
                    // row pitch fixed at P = 32 positions: the lane shifts never leave the warp (see the prob layer below)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        if (mt >= MT) continue;  // warp-uniform
                        float v[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float v1 = __shfl_down_sync(0xffffffffu, __uint_as_float(r[mt][8 + c]), 1);
                            const float v2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r[mt][16 + c]), 2);
                            v[c] = __uint_as_float(r[mt][c]) + v1 + v2 + s_shift[c];
                            if (L.relu) v[c] = fmaxf(v[c], 0.f);
                        }
                        if (!((vmask >> mt) & 1u)) continue;
                        uint4 pk;
                        pk.x = pack_act(v[0], v[1]);
                        pk.y = pack_act(v[2], v[3]);
                        pk.z = pack_act(v[4], v[5]);
                        pk.w = pack_act(v[6], v[7]);
                        reinterpret_cast<uint4 *>(L.out)[(size_t)b * plane + base[mt] + zoff] = pk;
                        if (L.out2 != nullptr)
                            reinterpret_cast<uint4 *>(L.out2)[(size_t)b * 4 * plane_sp + sp[mt] + (size_t)(zs + e) * zstride_sp] = pk;
                    }
                } else if constexpr (KWF) {
                    // prob layer (Cout = 1, kw folded into N): only columns 0..2 of a block are ever non-zero
                    uint32_t r[4][4];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
                        if (mt < MT) ptx::tmem_ld_x4(tb + mt * cols_mt, r[mt]);
                    ptx::tmem_ld_wait();
                    if (mirrored) {
                        uint32_t r2[4][4];
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt)
                            if (mt < MT) ptx::tmem_ld_x4(tb + R * CW + mt * cols_mt, r2[mt]);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                            for (int k = 0; k < 3; ++k) r[mt][k] = __float_as_uint(__uint_as_float(r[mt][k]) + __uint_as_float(r2[mt][k]));
                    }
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
                        if (mt < MT) {
                            ptx::tmem_st_zero_x4(tb + mt * cols_mt);
                            if (mirrored) ptx::tmem_st_zero_x4(tb + R * CW + mt * cols_mt);
                        }
                    ptx::tmem_st_wait();
                    ptx::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(tempty_bar(myblk));
                    // columns 0..2 are U_kw[p] = sum_{kh,ci} in[p + kh*P] w[kh,kw]; out[p] = U_0[p] + U_1[p+1] + U_2[p+2].
                    // The row pitch of this layer is fixed at P = 32 positions = one 32-lane group, and the last two positions
                    // of a row are halo columns, never outputs: p+1, p+2 are always lanes of the same warp (an earlier
                    // version with arbitrary P exchanged group edges through shared memory behind a named barrier of the
                    // set's 4 warps: 160 -> 152 us without it, despite 9 % more MMA rows).
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        if (mt >= MT) continue;  // warp-uniform
                        const float v1 = __shfl_down_sync(0xffffffffu, __uint_as_float(r[mt][1]), 1);
                        const float v2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r[mt][2]), 2);
                        if (!((vmask >> mt) & 1u)) continue;
                        float v = __uint_as_float(r[mt][0]) + v1 + v2 + s_shift[0];
                        if (L.relu) v = fmaxf(v, 0.f);
                        reinterpret_cast<float *>(L.out)[(size_t)b * plane + base[mt] + zoff] = v;
                    }
                } else {
                    // MB M-tiles at a time: 32 accumulator registers (+ 32 for a mirror block)
                    constexpr int MB = (CW == 32) ? 1 : 2;
#pragma unroll
                    for (int m0 = 0; m0 < 4; m0 += MB) {
                        if (m0 >= MT) break;  // warp-uniform
                        uint32_t r[MB][CW];
#pragma unroll
                        for (int i = 0; i < MB; ++i)
                            if (m0 + i < MT) {
#pragma unroll
                                for (int c8 = 0; c8 < CW / 8; ++c8)
                                    if (c8 < nchunk) ptx::tmem_ld_x8(tb + (m0 + i) * cols_mt + c8 * 8, &r[i][c8 * 8]);
                            }
                        ptx::tmem_ld_wait();
                        if (mirrored) {
                            uint32_t r2[MB][CW];
#pragma unroll
                            for (int i = 0; i < MB; ++i)
                                if (m0 + i < MT) {
#pragma unroll
                                    for (int c8 = 0; c8 < CW / 8; ++c8)
                                        if (c8 < nchunk) ptx::tmem_ld_x8(tb + R * CW + (m0 + i) * cols_mt + c8 * 8, &r2[i][c8 * 8]);
                                }
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < MB; ++i)
#pragma unroll
                                for (int k = 0; k < CW; ++k)
                                    if (k < 8 * nchunk) r[i][k] = __float_as_uint(__uint_as_float(r[i][k]) + __uint_as_float(r2[i][k]));
                        }
                        // values are in registers: zero the block (and its mirror) for its next use
#pragma unroll
                        for (int i = 0; i < MB; ++i)
                            if (m0 + i < MT) {
#pragma unroll
                                for (int c16 = 0; c16 < CW; c16 += 16) {
                                    ptx::tmem_st_zero_x16(tb + (m0 + i) * cols_mt + c16);
                                    if (mirrored) ptx::tmem_st_zero_x16(tb + R * CW + (m0 + i) * cols_mt + c16);
                                }
                            }
                        if (m0 + MB >= MT) {  // last batch: hand the block back
                            ptx::tmem_st_wait();
                            ptx::tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) ptx::mbar_arrive(tempty_bar(myblk));
                        }
#pragma unroll
                        for (int i = 0; i < MB; ++i) {
                            const int mt = m0 + i;
                            if (mt >= MT || !((vmask >> mt) & 1u)) continue;
                            const size_t vox = base[mt] + zoff;
#pragma unroll
                            for (int c8 = 0; c8 < CW / 8; ++c8) {
                                if (c8 >= nchunk) break;
                                float v[8];
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    v[k] = __uint_as_float(r[i][c8 * 8 + k]) + s_shift[c8 * 8 + k];
                                    if (L.relu) v[k] = fmaxf(v[k], 0.f);
                                }
                                if (L.out_f32) {
                                    reinterpret_cast<float *>(L.out)[(size_t)b * plane + vox] = v[0];
                                } else {
                                    uint4 pk;
                                    pk.x = pack_act(v[0], v[1]);
                                    pk.y = pack_act(v[2], v[3]);
                                    pk.z = pack_act(v[4], v[5]);
                                    pk.w = pack_act(v[6], v[7]);
                                    reinterpret_cast<uint4 *>(L.out)[((size_t)b * (L.cout_total / 8) + c8) * plane + vox] = pk;
                                    if (L.out2 != nullptr)
                                        reinterpret_cast<uint4 *>(L.out2)[((size_t)b * 4 * (L.cout_total >> 3) + c8) * plane_sp + sp[mt] +
                                                                          (size_t)(zs + e) * zstride_sp] = pk;
                                }
                            }
                        }
                    }
                }
                epi_work += clock64() - c1;
            }
            blk = (blk + 2) % R;  // the two lead-in blocks of the next item are never drained
        }
        if (warp == 2 && lane == 0 && L.dbg) {
            L.dbg[blockIdx.x * 12 + 6] = epi_wait;
            L.dbg[blockIdx.x * 12 + 7] = epi_work;
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------
// Layout conversion and weight packing kernels
// ------------------------------------------------------------------------------------------------
// fp32 NCDHW [B][C][N] -> bf16 CP8 [B][C/8][N][8]     (N = D*H*W voxels)
__global__ void ncdhw_to_cp8_kernel(const float *__restrict__ in, uint4 *__restrict__ out, int C, size_t N) {
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int bc = blockIdx.y;  // b * (C/8) + chunk
    if (v >= N) return;
    const int b = bc / (C / 8), chunk = bc % (C / 8);
    const float *src = in + ((size_t)b * C + chunk * 8) * N + v;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = __ldcs(src + (size_t)e * N);
    uint4 pk;
    pk.x = pack_act(f[0], f[1]);
    pk.y = pack_act(f[2], f[3]);
    pk.z = pack_act(f[4], f[5]);
    pk.w = pack_act(f[6], f[7]);
    out[(size_t)bc * N + v] = pk;
}

// bf16 CP8 -> fp32 NCDHW (tests / mixed-precision fallbacks)
__global__ void cp8_to_ncdhw_kernel(const uint4 *__restrict__ in, float *__restrict__ out, int C, size_t N) {
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int bc = blockIdx.y;
    if (v >= N) return;
    const int b = bc / (C / 8), chunk = bc % (C / 8);
    const uint4 pk = __ldg(in + (size_t)bc * N + v);
    const uint32_t *p = reinterpret_cast<const uint32_t *>(&pk);
    float *dst = out + ((size_t)b * C + chunk * 8) * N + v;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act(p[e]);
        dst[(size_t)(2 * e) * N] = f.x;
        dst[(size_t)(2 * e + 1) * N] = f.y;
    }
}

// One packed B block per op: [2 K-chunks][NPAD rows][8 bf16].  src describes where each chunk comes from.
struct WSrc {
    int16_t tap[2];   // filter tap index kd*9+kh*3+kw of each chunk, -1 = zero chunk
    int16_t cin0[2];  // first input channel of each chunk
};
struct WPackParams {
    int nblocks, npad, cout_group, cout_total, cin_total, ngroups, transposed;
    int fold_cw;  // > 0: depth-folded layout, B row n = (kd = 2 - n / fold_cw, cout = n % fold_cw); src taps are kh*3+kw
    int fold_kw;  // 1: depth-folded layout with kw in N too: B row n = (kd = 2 - n / 16, kw = n % 16 < 3), Cout = 1; src taps are kh*3
                  // 2: the same for Cout = 8 with 32-column blocks: B row n = (kd = 2 - n / 32, kw = (n % 32) / 8 < 3, cout = n % 8)
    int merged_t; // 1: class-merged transposed conv, B row n = (class n / cout, cout n % cout); src taps are dz*4+dy*2+dx
    int ntaps;    // taps per (cout, cin) pair in the source weights: 27 (3-D) or 9 (2-D, [Cout][Cin][3][3])
    int f16;      // 1: fp16 output, 0: bf16
    int kw2d;     // 1: 2-D layer with kw folded into N: B row n = (kw = n / cout < 3, cout = n % cout); src taps are kh*3
    WSrc src[kMaxOps];
};

__global__ void pack_weights_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ out,
                                    const __grid_constant__ WPackParams p) {
    const int per_group = p.nblocks * 2 * p.npad * 8;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_group * p.ngroups) return;
    const int g = idx / per_group;
    int r = idx % per_group;
    const int e = r % 8; r /= 8;
    const int n = r % p.npad; r /= p.npad;
    const int c = r % 2;
    const int blk = r / 2;
    int tap = p.src[blk].tap[c];
    const int cin = p.src[blk].cin0[c] + e;
    int co = g * p.cout_group + n;
    int nn = n;
    if (p.kw2d) {
        const int kw = n / p.cout_total;
        nn = n % p.cout_total;
        co = nn;
        if (kw >= 3) tap = -1;
        else if (tap >= 0) tap += kw;
    } else if (p.fold_cw > 0 && p.fold_kw == 2) {  // B row n = (kd = 2 - n / 32, kw = (n % 32) / 8 < 3, cout = n % 8)
        const int kw = (n % p.fold_cw) >> 3;
        nn = n & 7;
        co = nn;
        if (kw >= 3) tap = -1;
        else if (tap >= 0) tap += (2 - n / p.fold_cw) * 9 + kw;
    } else if (p.fold_cw > 0 && p.fold_kw) {
        const int kw = n % p.fold_cw;
        nn = 0;
        co = 0;
        if (kw >= 3) tap = -1;
        else if (tap >= 0) tap += (2 - n / p.fold_cw) * 9 + kw;
    } else if (p.fold_cw > 0) {
        nn = n % p.fold_cw;
        co = nn;
        if (tap >= 0) tap += (2 - n / p.fold_cw) * 9;
    }
    if (p.merged_t) {
        // output parity class a = (pz,py,px) uses input offset d = (dz,dy,dx) iff d <= p componentwise, through
        // the kernel tap k = 1 (p = 0), 2 (p = 1, d = 0) or 0 (p = 1, d = 1) per dimension
        const int a = n / p.cout_total;
        nn = n % p.cout_total;
        co = nn;
        if (tap >= 0) {
            const int pz = a >> 2, py = (a >> 1) & 1, px = a & 1, dz = tap >> 2, dy = (tap >> 1) & 1, dx = tap & 1;
            if (a >= 8 || dz > pz || dy > py || dx > px) tap = -1;
            else tap = (pz ? (dz ? 0 : 2) : 1) * 9 + (py ? (dy ? 0 : 2) : 1) * 3 + (px ? (dx ? 0 : 2) : 1);
        }
    }
    float v = 0.f;
    if (tap >= 0 && nn < p.cout_group && co < p.cout_total && cin < p.cin_total)
        v = p.transposed ? w[((size_t)cin * p.cout_total + co) * 27 + tap] : w[((size_t)co * p.cin_total + cin) * p.ntaps + tap];
    if (p.f16 || kActF16) reinterpret_cast<__half *>(out)[idx] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    else out[idx] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------
// Host side: per-layer configuration
// ------------------------------------------------------------------------------------------------
enum TcKind { TC_CONV_S1 = 0, TC_CONV_S2 = 1, TC_CONVT = 2, TC_CONV2D = 3 };

struct TcPlan {
    TcLayer L;
    WPackParams W;
    CUtensorMap tmap;
    int npad;
    bool kw2d;
    size_t smem_bytes;
    size_t wpacked_bytes;
    int grid;
};

static constexpr int kSmemLimit = 227 * 1024;

// Builds the plan for one layer.  in: bf16 CP8 [B][cin/8][Din][Hin][Win][8].
static int make_plan(TcPlan &pl, TcKind kind, int B, int cin, int cout, int Din, int Hin, int Win, const void *in_ptr,
                     int num_sms, bool encode = true, bool in_split = false, bool allow_kw2d = false,
                     const void *skip_ptr = nullptr) {
    TcLayer &L = pl.L;
    memset(&pl, 0, sizeof(pl));
    MVS_REQUIRE(cin % 8 == 0, "tc conv: Cin must be a multiple of 8");
    const int chunks = cin / 8;
    // tile space
    int Dt, Ht, Wt;
    if (kind == TC_CONV_S2) { Dt = Din / 2; Ht = Hin / 2; Wt = Win / 2; }
    else { Dt = Din; Ht = Hin; Wt = Win; }
    // N (padded cout) per CTA: keep resident weights <= ~112 KB
    const int kpairs_tap = (cin >= 16) ? cin / 16 : 1;
    // depth-folded variant (conv3d_tc_fold_kernel): stride-1 layers whose Cout fits one 16-column block
    // ... and Cout = 32 (conv4, 32 -> 32): 32-column blocks, N = 96: 18 MMAs of 56 cycles per plane and M-tile instead of 54 of 40
    const bool fold = (kind == TC_CONV_S1) && (cout <= 16 || (cout == 32 && cin >= 16));
    const bool fold_kw = fold && cin == 8 && cout == 1;
    // conv0 (32 -> 8): kw folded into N as well, 32-column blocks [kw][8 Cout] (24 used), N = 96: 6 MMAs of 56 cycles per
    // plane and M-tile instead of 18 of 44 -- the layer was bound by the shared-memory operand path of its MMAs
    const bool fold_kw8 = fold && cin >= 16 && cout == 8;
    const int fold_cw = fold ? ((fold_kw8 || cout == 32) ? 32 : 16) : 0;
    const bool is2d = (kind == TC_CONV2D);  // planes are independent images: only the (kh, kw) taps of one plane
    // class-merged transposed conv: one MMA per input offset (dz,dy,dx) and K-chunk with N = 8 classes x Cout
    const bool merged_t = (kind == TC_CONVT) && cin >= 16 && 8 * cout <= 128;
    // 2-D layers with fp16 operands (FeatureNet): kw folded into N (N = 3*Cout), row pitch fixed at 32 positions so that the
    // epilogue's lane shifts stay inside a warp
    // Measured per FeatureNet layer (tools/featurenet_tc_profile.py, kw-folded / plain, ms): 8->8 k3 0.119 / 0.099 (the
    // plain form has only 5 MMAs per M-tile: the fold's 3x larger accumulator read dominates), 32(s2d)->16 0.044 / 0.065,
    // 16->16 0.038 / 0.040, 64(s2d)->32 0.032 / 0.037, 32->32 0.027 / 0.021 -- so: folded when the plain form has >= 9
    // MMAs per M-tile and Cout <= 16, or >= 36 with Cout = 32.
    const bool kw2d = is2d && allow_kw2d  && Win >= 30 &&
                      (((cout == 8 || cout == 16) && cin >= 16) || (cout == 32 && cin >= 64));
    const bool skip_tma = merged_t && skip_ptr != nullptr;
    int ntaps_ops;  // MMA instructions per step
    if (merged_t) ntaps_ops = 8 * kpairs_tap;
    else if (fold_kw) ntaps_ops = 2;
    else if (kw2d) ntaps_ops = (cin >= 16) ? 3 * kpairs_tap : 2;
    else if (fold_kw8) ntaps_ops = 3 * kpairs_tap;
    else if (fold || is2d) ntaps_ops = (cin >= 16) ? 9 * kpairs_tap : 5;
    else if (cin >= 16) ntaps_ops = 27 * kpairs_tap;
    else ntaps_ops = (kind == TC_CONVT) ? 27 : 15;  // cin == 8: taps are paired (conv) / not paired (convT, unused)
    MVS_REQUIRE(!(kind == TC_CONVT && cin < 16), "tc convT needs Cin >= 16");
    MVS_REQUIRE(ntaps_ops <= kMaxOps, "tc conv: too many ops");
    int ngroups = 1;
    int cout_group = cout;
    while (!merged_t && !kw2d) {
        const int npad_try = fold ? 3 * fold_cw : (cout_group <= 16 ? 16 : (cout_group <= 32 ? 32 : 64));
        if ((size_t)ntaps_ops * npad_try * 32 <= 112 * 1024 && cout_group <= 64) break;
        ngroups *= 2;
        cout_group = cout / ngroups;
        MVS_REQUIRE(cout_group >= 8 && cout % ngroups == 0, "tc conv: cannot split Cout=%d", cout);
    }
    const int npad = kw2d ? std::max(32, 3 * cout) : merged_t ? 8 * cout : (fold ? 3 * fold_cw : (cout_group <= 16 ? 16 : (cout_group <= 32 ? 32 : 64)));
    const int nacc = (kind == TC_CONVT && !merged_t) ? 8 : 1;
    const int wbytes = ntaps_ops * npad * 32;
    // ring geometry
    const int need = (fold || is2d) ? 1 : ((kind == TC_CONVT) ? 2 : 3);
    const int adv = (kind == TC_CONV_S2) ? 2 : 1;
    const int nsub = (kind == TC_CONV_S2) ? 4 : 1;
    const int halo = (kind == TC_CONV_S1 || is2d) ? 2 : 1;  // extra rows / cols in a (sub-)plane box
    // choose the tile: TXB columns, TY rows, MT M-tiles of 128 flattened positions
    // boxDim[x] <= 256: 2 uint64 per voxel (merged inner dim) or elementStrides = 2; the skip tile box is 4*TXB uint64 wide
    const int max_cols = skip_tma ? 64 : 128 - halo;
    int best_TXB = 0, best_TY = 0, best_MT = 0, best_nslot = 0;
    double best_score = -1;
    // MT limit: 2 buffers x nacc x MT x npad <= 512 columns; folded: MT regions of R blocks x 16 columns, R >= 8
    const int npad_cols = kw2d ? (cout == 8 ? 32 : (cout == 16 ? 64 : 128)) : npad;  // TMEM columns per M-tile (the template's NPAD)
    const int tmem_budget = fold ? (fold_cw == 32 ? 2 : 4) : 256 / (nacc * npad_cols);
    for (int nx = 1; nx <= 64; ++nx) {
        const bool pitch32 = kw2d || fold_kw || fold_kw8;  // lane shifts of the kw-folded epilogues stay inside a warp
        const int TXB = pitch32 ? 30 : (Wt + nx - 1) / nx;
        if (pitch32 && nx > 1) break;
        if (TXB > max_cols) continue;
        const int P = TXB + halo;
        for (int MT = 1; MT <= 4 && MT <= tmem_budget; ++MT) {
            const int TY = std::min((MT * 128) / P, Ht);
            if (TY < 1) continue;
            const int rows = TY + halo;
            const size_t sub_bytes = (size_t)chunks * rows * P * 16;
            const size_t slot_bytes = nsub * ((sub_bytes + 127) & ~(size_t)127);
            for (int nslot = 8; nslot >= need; --nslot) {  // deeper ring = more TMA prefetch distance
                const size_t skip_smem = skip_tma ? 2 * (((size_t)128 * TXB * TY * (cout / 8) + 1023) & ~(size_t)1023) + 1024 : 0;
                const size_t total = 256 + kMaxOps * 8 + 256 + wbytes + 128 + nslot * slot_bytes + (128 + 2 * P + 8) * 16 + 1024 + 768 +
                                     (fold ? 512 : 0) + skip_smem;
                if (total > (size_t)kSmemLimit) continue;
                // useful fraction of the MMA rows, x- and y-tile padding, halo re-read
                const double useful = (double)(TY * TXB) / (MT * 128.0);
                const double xeff = pitch32 ? (double)Wt / (((Wt + TXB - 1) / TXB) * TXB) : (double)Wt / (nx * TXB);
                const double yeff = (double)Ht / (((Ht + TY - 1) / TY) * TY);
                const double pipe = (nslot >= need + 2 * adv) ? 1.0 : (nslot >= need + adv ? 0.9 : 0.6);
                const double score = useful * xeff * yeff * pipe * (0.9 + 0.1 * (double)TY / rows);
                if (score > best_score) {
                    best_score = score; best_TXB = TXB; best_TY = TY; best_MT = MT; best_nslot = nslot;
                }
                break;
            }
        }
        if (nx > 1 && best_score > 0 && (Wt + nx - 1) / nx < 24) break;
    }
    MVS_REQUIRE(best_score > 0, "tc conv: no tile configuration fits shared memory (cin=%d cout=%d)", cin, cout);
    const int TXB = best_TXB, TY = best_TY, MT = best_MT, P = TXB + halo, rows = TY + halo;

    L.B = B; L.Dt = Dt; L.Ht = Ht; L.Wt = Wt;
    L.TXB = TXB; L.TY = TY; L.P = P; L.MT = MT;
    L.tiles_x = (Wt + TXB - 1) / TXB;
    L.tiles_y = (Ht + TY - 1) / TY;
    L.ngroups = ngroups;
    L.nsub = nsub; L.chunks = chunks;
    L.sub_bytes = chunks * rows * P * 16;
    L.sub_stride = (L.sub_bytes + 127) & ~127;
    L.slot_bytes = nsub * L.sub_stride;
    L.need = need; L.adv = adv; L.nslot = best_nslot;
    L.pz0 = (kind == TC_CONVT || is2d) ? 0 : -1;
    L.zscale = (kind == TC_CONV_S2) ? 2 : 1;
    L.in_scale = (kind == TC_CONV_S2) ? 2 : 1;
    if (kind == TC_CONV_S1 || is2d) { L.sub_xoff[0] = -1; L.sub_yoff[0] = -1; }
    else if (kind == TC_CONVT) { L.sub_xoff[0] = 0; L.sub_yoff[0] = 0; }
    else {
        for (int s = 0; s < 4; ++s) {  // s = ypar*2 + xpar; parity 1 = odd input index, starts one element earlier
            L.sub_yoff[s] = (s >> 1) ? -1 : 0;
            L.sub_xoff[s] = (s & 1) ? -1 : 0;
        }
    }
    // z segmentation: enough work items for ~2 per SM
    const int cols = B * L.tiles_x * L.tiles_y * ngroups;
    int zsegs = 1;
    {
        double best = -1;
        for (int zs = 1; zs <= (is2d ? Dt : std::max(1, Dt / 4)); ++zs) {
            const int len = (Dt + zs - 1) / zs, nseg = (Dt + len - 1) / len;
            const long long items = (long long)cols * nseg;
            const double wave = (double)items / ((double)((items + num_sms - 1) / num_sms) * num_sms);
            const double haloeff = fold ? (double)len / (len + 2) : (double)(adv * len) / (adv * (len - 1) + need);
            const double sc = wave * haloeff * (is2d ? (double)len / (len + 0.05) : 1.0);  // 2-D: per-item setup

            if (sc > best + 1e-9) { best = sc; zsegs = zs; }
        }
    }
    L.zseg_len = (Dt + zsegs - 1) / zsegs;
    L.zsegs = (Dt + L.zseg_len - 1) / L.zseg_len;
    L.n_items = cols * L.zsegs;
    MVS_REQUIRE(L.n_items < (1 << 20), "tc conv: too many work items (%d)", L.n_items);
    auto rcp40 = [](int d) { return (unsigned long long)(((1ull << 40) + (unsigned long long)d - 1) / (unsigned long long)d); };
    L.rcp_zsegs = rcp40(L.zsegs); L.rcp_tiles_x = rcp40(L.tiles_x); L.rcp_tiles_y = rcp40(L.tiles_y);
    L.nacc = nacc; L.npad = npad; L.wbytes_group = wbytes;
    L.tmem_bufs_log2 = (!fold && 4 * nacc * MT * npad_cols <= 512) ? 2 : 1;
    L.cout_group = cout_group; L.cout_total = cout;
    L.out_scale = (kind == TC_CONVT) ? 2 : 1;
    L.Dout = L.out_scale * Dt; L.Hout = L.out_scale * Ht; L.Wout = L.out_scale * Wt;

    // ---- op table + weight sources
    WPackParams &W = pl.W;
    W.npad = npad; W.cout_group = cout_group; W.cout_total = cout; W.cin_total = cin; W.ngroups = ngroups;
    W.transposed = (kind == TC_CONVT);
    W.fold_cw = fold_cw;
    W.fold_kw = fold_kw ? 1 : (fold_kw8 ? 2 : 0);
    L.fold_kw = W.fold_kw;
    W.ntaps = is2d ? 9 : 27;
    W.kw2d = kw2d ? 1 : 0;
    L.mma_n = kw2d ? npad : 0;
    W.merged_t = merged_t ? 1 : 0;
    L.merged_t = merged_t ? 1 : 0;
    L.fold = fold ? 1 : 0;
    L.dual = (!fold) ? 1 : 0;
    L.fold_R = fold ? std::min(kFoldMaxR, 512 / (MT * fold_cw)) : 0;
    L.fold_sets = fold ? ((fold_kw8 || fold_kw) ? 3 : 2) : 0;  // the plain variant needs > 128 registers: 2 sets
    const int chunk_stride = rows * P * 16;
    int nops = 0;
    auto tap_off = [&](int kh, int kw) -> int {  // byte offset of a conv tap inside its plane (chunk 0)
        if (kind == TC_CONV_S1 || is2d) return (kh * P + kw) * 16;
        const int ypar = (kh != 1), xpar = (kw != 1);
        return (ypar * 2 + xpar) * L.sub_stride + ((kh == 2 ? 1 : 0) * P + (kw == 2 ? 1 : 0)) * 16;
    };
    if (kind != TC_CONVT) {
        L.acc_first[0] = 0;
        for (int kd = 0; kd < ((fold || is2d) ? 1 : 3); ++kd) {  // folded: kd lives in the rows of the packed B block; 2-D: no kd
            if (cin >= 16) {
                for (int kh = 0; kh < 3; ++kh)
                    for (int kw = 0; kw < ((fold_kw8 || kw2d) ? 1 : 3); ++kw)  // kw folds: kw lives in the rows of the packed B block
                        for (int kc = 0; kc < cin / 16; ++kc) {
                            TcOp &op = L.ops[nops];
                            op.a_off = tap_off(kh, kw) + 2 * kc * chunk_stride;
                            op.lbo = chunk_stride;
                            op.widx = nops; op.plane_rel = kd; op.acc = 0;
                            const int tap = kd * 9 + kh * 3 + kw;
                            W.src[nops] = WSrc{{(int16_t)tap, (int16_t)tap}, {(int16_t)(16 * kc), (int16_t)(16 * kc + 8)}};
                            ++nops;
                        }
            } else if (fold_kw || kw2d) {  // K = (kh, 8 channels): kh = 0,1 in one instruction, kh = 2 (+ a zero chunk) in the other
                for (int i = 0; i < 2; ++i) {
                    TcOp &op = L.ops[nops];
                    op.a_off = tap_off(2 * i, 0);
                    op.lbo = (i == 0) ? tap_off(1, 0) - tap_off(0, 0) : 0;
                    op.widx = nops; op.plane_rel = kd; op.acc = 0;
                    W.src[nops] = WSrc{{(int16_t)(6 * i), (int16_t)(i == 0 ? 3 : -1)}, {0, 0}};
                    ++nops;
                }
            } else {  // cin == 8: one K=16 instruction covers two taps of the same plane (sorted by offset)
                int order[9];
                for (int i = 0; i < 9; ++i) order[i] = i;
                std::sort(order, order + 9, [&](int a, int b) { return tap_off(a / 3, a % 3) < tap_off(b / 3, b % 3); });
                for (int i = 0; i < 9; i += 2) {
                    TcOp &op = L.ops[nops];
                    const int t0 = order[i], t1 = (i + 1 < 9) ? order[i + 1] : -1;
                    op.a_off = tap_off(t0 / 3, t0 % 3);
                    op.lbo = (t1 >= 0) ? tap_off(t1 / 3, t1 % 3) - op.a_off : 0;
                    op.widx = nops; op.plane_rel = kd; op.acc = 0;
                    W.src[nops] = WSrc{{(int16_t)(kd * 9 + t0), (int16_t)(t1 >= 0 ? kd * 9 + t1 : -1)}, {0, 0}};
                    ++nops;
                }
            }
        }
        L.acc_first[1] = nops;
        L.acc_pz[0] = L.acc_py[0] = L.acc_px[0] = 0;
    } else if (merged_t) {
        // one op per input offset d = (dz,dy,dx) and K-chunk; the packed B block holds, for every parity class, the
        // tap that class applies to this offset (or zeros), see pack_weights_kernel
        L.acc_first[0] = 0;
        for (int d = 0; d < 8; ++d) {
            const int dz = d >> 2, dy = (d >> 1) & 1, dx = d & 1;
            for (int kc = 0; kc < cin / 16; ++kc) {
                TcOp &op = L.ops[nops];
                op.a_off = (dy * P + dx) * 16 + 2 * kc * chunk_stride;
                op.lbo = chunk_stride;
                op.widx = nops; op.plane_rel = dz; op.acc = 0;
                W.src[nops] = WSrc{{(int16_t)d, (int16_t)d}, {(int16_t)(16 * kc), (int16_t)(16 * kc + 8)}};
                ++nops;
            }
        }
        L.acc_first[1] = nops;
        for (int a = 0; a < 8; ++a) { L.acc_pz[a] = a >> 2; L.acc_py[a] = (a >> 1) & 1; L.acc_px[a] = a & 1; }
    } else {
        // 8 output-parity classes; per dimension: parity 0 -> (k=1, d=0); parity 1 -> (k=2, d=0), (k=0, d=1)
        for (int a = 0; a < 8; ++a) {
            const int pz = a >> 2, py = (a >> 1) & 1, px = a & 1;
            L.acc_first[a] = nops;
            L.acc_pz[a] = pz; L.acc_py[a] = py; L.acc_px[a] = px;
            for (int dz = 0; dz <= pz; ++dz)
                for (int dy = 0; dy <= py; ++dy)
                    for (int dx = 0; dx <= px; ++dx) {
                        const int kd = pz ? (dz ? 0 : 2) : 1, kh = py ? (dy ? 0 : 2) : 1, kw = px ? (dx ? 0 : 2) : 1;
                        for (int kc = 0; kc < cin / 16; ++kc) {
                            TcOp &op = L.ops[nops];
                            op.a_off = (dy * P + dx) * 16 + 2 * kc * chunk_stride;
                            op.lbo = chunk_stride;
                            op.widx = nops; op.plane_rel = dz; op.acc = a;
                            const int tap = kd * 9 + kh * 3 + kw;
                            W.src[nops] = WSrc{{(int16_t)tap, (int16_t)tap}, {(int16_t)(16 * kc), (int16_t)(16 * kc + 8)}};
                            ++nops;
                        }
                    }
        }
        L.acc_first[8] = nops;
    }
    MVS_REQUIRE(nops == ntaps_ops, "tc conv: internal op count mismatch (%d vs %d)", nops, ntaps_ops);
    for (int a = 0; a < (merged_t ? 8 : nacc); ++a) L.acc_voff[a] = (L.acc_pz[a] * L.Hout + L.acc_py[a]) * L.Wout + L.acc_px[a];
    MVS_REQUIRE((long long)L.Dout * L.Hout * L.Wout < (1LL << 31), "tc conv: volume too large for 32-bit class offsets");
    L.nops = nops;
    W.nblocks = nops;
    {   // order ops by (accumulator group, ring plane); the packed weight block follows its op (widx = position)
        std::vector<int> order(nops);
        for (int i = 0; i < nops; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            if (L.ops[a].acc != L.ops[b].acc) return L.ops[a].acc < L.ops[b].acc;
            return L.ops[a].plane_rel < L.ops[b].plane_rel;
        });
        std::vector<TcOp> ops2(nops);
        std::vector<WSrc> src2(nops);
        for (int i = 0; i < nops; ++i) { ops2[i] = L.ops[order[i]]; src2[i] = W.src[order[i]]; ops2[i].widx = (uint16_t)i; }
        for (int i = 0; i < nops; ++i) { L.ops[i] = ops2[i]; W.src[i] = src2[i]; }
        L.nseg = 0;
        for (int i = 0; i < nops; ++i) {
            const bool new_acc = (i == 0) || L.ops[i].acc != L.ops[i - 1].acc;
            if (new_acc || L.ops[i].plane_rel != L.ops[i - 1].plane_rel) {
                const int sg = L.nseg++;
                L.seg_first[sg] = (uint16_t)i; L.seg_plane[sg] = L.ops[i].plane_rel; L.seg_acc[sg] = L.ops[i].acc;
                L.seg_new_acc[sg] = new_acc;
                if (sg > 0) L.seg_last[sg - 1] = (uint16_t)i;
            }
        }
        L.seg_last[L.nseg - 1] = (uint16_t)nops;
    }
    for (int o = 0; o < nops; ++o) {
        TcOp &op = L.ops[o];
        op.a_lo = ((op.a_off >> 4) & 0x3FFFu) | ((op.lbo >> 4) << 16);
        op.b_lo = (((uint32_t)op.widx * npad * 32) >> 4) | (((uint32_t)npad * 16 >> 4) << 16);
    }
    pl.npad = npad_cols;
    pl.kw2d = kw2d;
    pl.wpacked_bytes = (size_t)ngroups * wbytes;
    L.skip_tma = skip_tma ? 1 : 0;
    L.skip_tx_bytes = 128 * TXB * TY * (cout / 8);
    L.skip_buf_bytes = (L.skip_tx_bytes + 1023) & ~1023;
    L.skip_off = (128 + 2 * P + 8) * 16;  // past the A-operand overrun guard that follows the ring
    pl.smem_bytes = 256 + kMaxOps * 8 + 256 + wbytes + 128 + (size_t)L.nslot * L.slot_bytes + (128 + 2 * P + 8) * 16 + 1024 + 768 +
                    (fold ? 512 : 0) + (skip_tma ? 2 * (size_t)L.skip_buf_bytes + 1024 : 0);
    pl.grid = std::min(L.n_items, num_sms);
    pl.grid = std::max(ngroups, pl.grid / ngroups * ngroups);  // every group gets the same number of CTAs

    if (!encode) return MVS_OK;
    // ---- tensor map over the input: dims (8ch, x, y, z, B*chunks)
    tmap_encode_fn enc = get_tmap_encode();
    MVS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    if (skip_tma) {  // skip tensor = bf16 CP8 [B][Cout/8][Dout][Hout][Wout][8], merged (8ch, x) inner dimension
        const int cpc = cout / 8;
        cuuint64_t gd[4] = {(cuuint64_t)2 * L.Wout, (cuuint64_t)L.Hout, (cuuint64_t)L.Dout, (cuuint64_t)B * cpc};
        cuuint64_t gs[3] = {(cuuint64_t)L.Wout * 16, (cuuint64_t)L.Wout * L.Hout * 16, (cuuint64_t)L.Wout * L.Hout * L.Dout * 16};
        cuuint32_t bx[4] = {(cuuint32_t)(4 * TXB), (cuuint32_t)(2 * TY), 2, (cuuint32_t)cpc};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult cs = enc(&L.skip_map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void *>(skip_ptr), gd, gs, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cs != CUDA_SUCCESS) return set_error(MVS_ERR_CUDA, "cuTensorMapEncodeTiled (skip tile) failed (%d)", (int)cs);
    }
    L.merged_x = (kind != TC_CONV_S2);
    L.in_split = (kind == TC_CONV_S2 && in_split) ? 1 : 0;
    if (L.in_split) {
        // parity-split input [B][4 (ypar,xpar)][chunks][D][H/2][W/2][8]: every sub-plane box is unit-stride, rows contiguous
        const int W2 = Win / 2, H2 = Hin / 2;
        cuuint64_t gdim4[4] = {(cuuint64_t)2 * W2, (cuuint64_t)H2, (cuuint64_t)Din, (cuuint64_t)B * 4 * chunks};
        cuuint64_t gstr4[3] = {(cuuint64_t)W2 * 16, (cuuint64_t)W2 * H2 * 16, (cuuint64_t)W2 * H2 * Din * 16};
        cuuint32_t box4[4] = {(cuuint32_t)(2 * P), (cuuint32_t)rows, 1, (cuuint32_t)chunks};
        cuuint32_t estr4[4] = {1, 1, 1, 1};
        CUresult cr4 = enc(&pl.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void *>(in_ptr), gdim4, gstr4, box4,
                           estr4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr4 != CUDA_SUCCESS) return set_error(MVS_ERR_CUDA, "cuTensorMapEncodeTiled (parity-split) failed (%d)", (int)cr4);
        return MVS_OK;
    }
    if (L.merged_x) {
        // A TMA request per 16-byte inner row is ~10 cycles; merging (8ch, x) into one contiguous inner
        // dimension of uint64 elements makes every box row one 16*P-byte request.
        cuuint64_t gdim4[4] = {(cuuint64_t)2 * Win, (cuuint64_t)Hin, (cuuint64_t)Din, (cuuint64_t)B * chunks};
        cuuint64_t gstr4[3] = {(cuuint64_t)Win * 16, (cuuint64_t)Win * Hin * 16, (cuuint64_t)Win * Hin * Din * 16};
        cuuint32_t box4[4] = {(cuuint32_t)(2 * P), (cuuint32_t)rows, 1, (cuuint32_t)chunks};
        cuuint32_t estr4[4] = {1, 1, 1, 1};
        CUresult cr4 = enc(&pl.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void *>(in_ptr), gdim4, gstr4, box4,
                           estr4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr4 != CUDA_SUCCESS) return set_error(MVS_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed (%d)", (int)cr4);
        return MVS_OK;
    }
    const int es = (kind == TC_CONV_S2) ? 2 : 1;
    cuuint64_t gdim[5] = {8, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)Din, (cuuint64_t)B * chunks};
    cuuint64_t gstr[4] = {16, (cuuint64_t)Win * 16, (cuuint64_t)Win * Hin * 16, (cuuint64_t)Win * Hin * Din * 16};
    cuuint32_t box[5] = {8, (cuuint32_t)(P * es), (cuuint32_t)(rows * es), 1, (cuuint32_t)chunks};
    cuuint32_t estr[5] = {1, (cuuint32_t)es, (cuuint32_t)es, 1, 1};
    CUresult cr = enc(&pl.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(in_ptr), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return set_error(MVS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return MVS_OK;
}

static long long *g_tc_dbg = nullptr;  // set by mvs_tc_set_debug_buffer (diagnostics only)

// ------------------------------------------------------------------------------------------------
// Packed-weight cache.  Packing (fp32 -> bf16/fp16 B-operand blocks in op order) depends only on the weights and the
// layer's op table, not on the activations, so it is done once per (device, weight pointer, layer configuration) and
// reused by later calls: 27 tiny launches per depth map disappear from the step.  The cache is keyed by POINTER: a
// caller that changes weights in place or frees and re-creates them must call mvs_weight_cache_clear() (the Python
// host does whenever it re-folds BatchNorm).  An event orders first use on one stream before reuse on another.
// Lifetime: a forward pass (costreg_tc / featurenet_tc) holds g_wcache_life SHARED from its first lookup until its last
// kernel is enqueued; a clear takes it EXCLUSIVELY, so every launch that uses an entry is already in a stream when the
// entry is freed, and cudaFree waits for the device.  (No second-chance "graveyard": two clears from other threads could
// free a pointer a third thread had looked up but not launched on yet.)
// ------------------------------------------------------------------------------------------------
struct WCacheKey {
    int dev;
    const void *w;
    uint64_t cfg;  // hash of WPackParams (op order, padding, dtype) or a caller tag
    bool operator<(const WCacheKey &o) const {
        if (dev != o.dev) return dev < o.dev;
        if (w != o.w) return w < o.w;
        return cfg < o.cfg;
    }
};
struct WCacheEntry {
    void *ptr;
    cudaEvent_t ready;
    bool done;  // the fill has been observed complete: no event wait needed any more
};
static std::mutex g_wcache_mu;            // the map itself
static std::shared_mutex g_wcache_life;   // entry lifetime, see above
static std::map<WCacheKey, WCacheEntry> g_wcache;

static uint64_t fnv1a(const void *data, size_t n, uint64_t h = 1469598103934665603ull) {
    const uint8_t *p = (const uint8_t *)data;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

// Returns the cached device buffer for `key`, creating it with fill(dst) (which must enqueue work on st) on a miss.
template <typename F>
static int wcache_get(const WCacheKey &key, size_t bytes, cudaStream_t st, void **out, F fill, bool *settled = nullptr) {
    std::lock_guard<std::mutex> lock(g_wcache_mu);
    auto it = g_wcache.find(key);
    if (settled) *settled = false;
    if (it == g_wcache.end()) {
        WCacheEntry e;
        e.done = false;
        MVS_CUDA(cudaMalloc(&e.ptr, bytes));
        if (cudaError_t ce = cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming); ce != cudaSuccess) {
            cudaFree(e.ptr);
            return set_error(MVS_ERR_CUDA, "cudaEventCreate (weight cache) failed: %s", cudaGetErrorString(ce));
        }
        if (int rc = fill(e.ptr)) { cudaFree(e.ptr); cudaEventDestroy(e.ready); return rc; }
        MVS_CUDA(cudaEventRecord(e.ready, st));
        it = g_wcache.emplace(key, e).first;
    } else if (it->second.done) {
        if (settled) *settled = true;
    } else if (cudaEventQuery(it->second.ready) == cudaSuccess) {
        it->second.done = true;  // filled long ago: nothing to order against (and no stream op between two conv launches)
        if (settled) *settled = true;
    } else {
        (void)cudaGetLastError();  // cudaErrorNotReady is not sticky, but leave no stale status behind
        MVS_CUDA(cudaStreamWaitEvent(st, it->second.ready, 0));
    }
    *out = it->second.ptr;
    return MVS_OK;
}

int weight_cache_clear() {
    std::unique_lock<std::shared_mutex> life(g_wcache_life);  // no forward pass is between a lookup and its launches
    std::lock_guard<std::mutex> lock(g_wcache_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto &kv : g_wcache) {
        cudaSetDevice(kv.first.dev);
        cudaFree(kv.second.ptr);  // waits for every kernel already enqueued on the device
        cudaEventDestroy(kv.second.ready);
    }
    g_wcache.clear();
    cudaSetDevice(cur);
    return MVS_OK;
}

static int run_layer(TcKind kind, const void *in, const float *w_fp32, const float *shift, int relu, const void *skip,
                     void *out, int out_f32, void *wpacked_scratch, int B, int cin, int cout, int Din, int Hin, int Win,
                     int num_sms, cudaStream_t st, int f16 = 0, int out_mode = 0, bool cache_weights = false, void *out2 = nullptr,
                     bool in_split = false) {
    // Plans (tile geometry, op table, tensor maps) depend only on the layer's shape and on the addresses baked into the
    // tensor maps.  The workspaces of consecutive forward passes are the same buffers, so a small per-thread cache turns the
    // per-launch planning and cuTensorMapEncodeTiled calls (~10 us per layer, 20 layers per depth map) into a lookup.
    struct PlanKey {
        int kind, B, cin, cout, Din, Hin, Win, num_sms, in_split, allow_kw2d;
        const void *in, *skip;
        bool operator==(const PlanKey &o) const { return memcmp(this, &o, sizeof(PlanKey)) == 0; }
    };
    struct PlanSlot { PlanKey key; TcPlan plan; bool used; };
    constexpr int kPlanSlots = 48;
    static thread_local std::vector<PlanSlot> plan_cache;
    static thread_local int plan_next = 0;
    PlanKey key;
    memset(&key, 0, sizeof(key));  // padding bytes too: the comparison is bytewise
    key.kind = (int)kind; key.B = B; key.cin = cin; key.cout = cout; key.Din = Din; key.Hin = Hin; key.Win = Win;
    key.num_sms = num_sms; key.in_split = in_split ? 1 : 0; key.allow_kw2d = (f16 != 0 && skip == nullptr && !out_f32) ? 1 : 0;
    key.in = in; key.skip = skip;
    TcPlan *plp = nullptr;
    for (auto &sl : plan_cache)
        if (sl.used && sl.key == key) { plp = &sl.plan; break; }
    if (plp == nullptr) {
        if ((int)plan_cache.size() < kPlanSlots) plan_cache.resize(plan_cache.size() + 1), plp = &plan_cache.back().plan, plan_cache.back().used = false;
        else { plan_next = (plan_next + 1) % kPlanSlots; plan_cache[plan_next].used = false; plp = &plan_cache[plan_next].plan; }
        PlanSlot *slot = reinterpret_cast<PlanSlot *>(reinterpret_cast<char *>(plp) - offsetof(PlanSlot, plan));
        if (int rc = make_plan(*plp, kind, B, cin, cout, Din, Hin, Win, in, num_sms, true, in_split, key.allow_kw2d != 0, skip)) return rc;
        slot->key = key;
        slot->used = true;
    }
    TcPlan &pl = *plp;
    MVS_REQUIRE(out2 == nullptr || (out_mode == 0 && !out_f32 && !f16 && (pl.L.fold || pl.L.nacc == 1) && skip == nullptr && Hin % 2 == 0 && Win % 2 == 0), "second output: plain conv layers only");
    pl.L.out2 = out2;
    MVS_REQUIRE(!f16 || (kind == TC_CONV2D && (pl.npad <= 32 || pl.kw2d) && skip == nullptr && !out_f32), "fp16 operands: 2-D layers only");
    MVS_REQUIRE(out_mode == 0 || (kind == TC_CONV2D && B == 1), "alternative output layouts: 2-D layers only");
    MVS_REQUIRE(out_mode != 1 || (Hin % 2 == 0 && Win % 2 == 0), "space-to-depth output needs even H, W");
    pl.L.f16 = f16;
    pl.L.out_mode = out_mode;
    pl.W.f16 = f16;
    pl.L.out = out;
    pl.L.skip = (const uint4 *)skip;
    pl.L.shift = shift;
    pl.L.relu = relu;
    pl.L.out_f32 = out_f32;
    pl.L.dbg = g_tc_dbg;
    if (!cache_weights) {  // single-layer entry points: ad-hoc weight tensors, pack into the caller's scratch every call
        const int nw = (int)(pl.wpacked_bytes / 2);
        pack_weights_kernel<<<cdiv(nw, 256), 256, 0, st>>>(w_fp32, (__nv_bfloat16 *)wpacked_scratch, pl.W);
        MVS_LAUNCH_CHECK(1);
        pl.L.wpacked = (const uint4 *)wpacked_scratch;
    } else {
        int dev = 0;
        MVS_CUDA(cudaGetDevice(&dev));
        const int nw = (int)(pl.wpacked_bytes / 2);
        void *wp = nullptr;
        bool settled = false;
        const WCacheKey key{dev, w_fp32, fnv1a(&pl.W, sizeof(pl.W))};
        if (int rc = wcache_get(key, pl.wpacked_bytes, st, &wp, [&](void *dst) -> int {
                pack_weights_kernel<<<cdiv(nw, 256), 256, 0, st>>>(w_fp32, (__nv_bfloat16 *)dst, pl.W);
                MVS_LAUNCH_CHECK(1);
                return MVS_OK;
            }, &settled))
            return rc;
        pl.L.wpacked = (const uint4 *)wp;
        pl.L.w_early = settled ? 1 : 0;
    }
    auto launch = [&](auto kern) -> int {
        // a per-function attribute shared by every host thread: always the same value (the opt-in maximum), never a
        // per-layer one that a concurrent launch of another layer could lower between this call and the launch
        // once per (kernel, device) and host thread.  Every kernel has the same function-pointer TYPE, so this lambda has
        // one instantiation: the table is keyed by the pointer's value
        struct AttrDone { const void *fn; uint64_t devs; };
        static thread_local AttrDone attr_done[32] = {};
        int adev = 0;
        MVS_CUDA(cudaGetDevice(&adev));
        AttrDone *ad = nullptr;
        for (auto &a : attr_done)
            if (a.fn == (const void *)kern || a.fn == nullptr) { ad = &a; break; }
        if (ad == nullptr || adev >= 64 || ad->fn == nullptr || !((ad->devs >> adev) & 1ull)) {
            MVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
            if (ad != nullptr && adev < 64) { ad->fn = (const void *)kern; ad->devs |= 1ull << adev; }
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)pl.grid);
        cfg.blockDim = dim3((unsigned)(pl.L.fold ? fold_threads(pl.L.fold_sets) : (pl.L.dual ? kTcThreadsDual : kTcThreads)));
        cfg.dynamicSmemBytes = pl.smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MVS_CUDA(cudaLaunchKernelEx(&cfg, kern, pl.tmap, pl.L));
        MVS_LAUNCH_CHECK(1);
        return MVS_OK;
    };
    if (pl.L.fold) {
        if (pl.L.fold_kw == 2) return launch(conv3d_tc_fold_kernel<32, 3, true>);
        if (pl.L.fold_kw) return launch(conv3d_tc_fold_kernel<16, 3, true>);
        return pl.W.fold_cw == 32 ? launch(conv3d_tc_fold_kernel<32, 2, false>) : launch(conv3d_tc_fold_kernel<16, 2, false>);
    }
    // lean epilogue when there is one accumulator per M-tile, no skip connection and a 16-bit output
    const bool simple = (pl.L.nacc == 1) && (skip == nullptr) && !out_f32 && !pl.L.merged_t;
    if (pl.kw2d) {
        if (pl.npad == 32) return launch(conv3d_tc_kernel<32, true, 3>);
        return pl.npad == 64 ? launch(conv3d_tc_kernel<64, true, 3>) : launch(conv3d_tc_kernel<128, true, 3>);
    }
    if (f16) return pl.npad == 16 ? launch(conv3d_tc_kernel<16, true, 1>) : launch(conv3d_tc_kernel<32, true, 1>);
    if (pl.L.merged_t) return pl.npad == 64 ? launch(conv3d_tc_kernel<64, false, 2>) : launch(conv3d_tc_kernel<128, false, 2>);
    if (simple) {
        if (pl.npad == 16) return launch(conv3d_tc_kernel<16, false, 1>);
        if (pl.npad == 32) return launch(conv3d_tc_kernel<32, false, 1>);
        return launch(conv3d_tc_kernel<64, false, 1>);
    }
    if (pl.npad == 16) return launch(conv3d_tc_kernel<16>);
    if (pl.npad == 32) return launch(conv3d_tc_kernel<32>);
    return launch(conv3d_tc_kernel<64>);
}

// ------------------------------------------------------------------------------------------------
// CostRegNet.forward (mvsnet.py:64-73) on tensor cores.
// workspace (bytes): bf16 CP8 activations  vol 64 N0 | c0 16 | c1 4 | c2 4 | c3 1 | c4 1 | c5 .25 | c6 .25 |
//                    u7 1 | u9 4 | u11 16   (N0 = B*D*H*W voxels; bytes per voxel shown) + packed weights
// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t n, size_t a) { return (n + a - 1) / a * a; }
static constexpr size_t kWScratch = 512 * 1024;

size_t costreg_tc_workspace_bytes(int B, int D, int H, int W) {
    const size_t n0 = (size_t)B * D * H * W;
    const size_t act = n0 * 64 + n0 * 16 + n0 * 4 + n0 * 4 + n0 + n0 + n0 / 4 + n0 / 4 + n0 + n0 * 4 + n0 * 16 +
                       n0 * 16 + n0 * 4 + n0;  // + the parity-split copies of c0, c2, c4 read by the stride-2 layers
    return align_up(act, 1024) + 11 * kWScratch + 16 * 1024;
}

int costreg_tc(const float *volume, const void *volume_cp8, const mvs_costreg_params *p, float *logits, void *workspace,
               int B, int D, int H, int W, cudaStream_t st) {
    std::shared_lock<std::shared_mutex> lease(g_wcache_life);  // cached weight pointers stay valid until the last launch
    int dev = 0, num_sms = 148;
    MVS_CUDA(cudaGetDevice(&dev));
    MVS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t n0 = (size_t)B * D * H * W;
    uint8_t *ws = (uint8_t *)workspace;
    auto take = [&](size_t bytes) { uint8_t *r = ws; ws += align_up(bytes, 1024); return r; };
    void *vol = take(volume_cp8 ? 0 : n0 * 64), *c0 = take(n0 * 16), *c1 = take(n0 * 4), *c2 = take(n0 * 4), *c3 = take(n0),
         *c4 = take(n0), *c5 = take(n0 / 4), *c6 = take(n0 / 4), *u7 = take(n0), *u9 = take(n0 * 4),
         *u11 = take(n0 * 16);
    void *c0s = take(n0 * 16), *c2s = take(n0 * 4), *c4s = take(n0);
    uint8_t *wsc = take(11 * kWScratch);
    const size_t N = (size_t)D * H * W;
    if (volume_cp8) {
        vol = const_cast<void *>(volume_cp8);
    } else {
        ncdhw_to_cp8_kernel<<<dim3(cdiv(N, 256), B * 4), 256, 0, st>>>(volume, (uint4 *)vol, 32, N);
        MVS_LAUNCH_CHECK(1);
    }
    int rc;
#define RUN(expr) if ((rc = (expr)) != MVS_OK) return rc
    RUN(run_layer(TC_CONV_S1, vol, p->w[0], p->shift[0], 1, nullptr, c0, 0, wsc + 0 * kWScratch, B, 32, 8, D, H, W, num_sms, st, 0, 0, true, c0s));
    RUN(run_layer(TC_CONV_S2, c0s ? c0s : c0, p->w[1], p->shift[1], 1, nullptr, c1, 0, wsc + 1 * kWScratch, B, 8, 16, D, H, W, num_sms, st, 0, 0, true, nullptr, c0s != nullptr));
    RUN(run_layer(TC_CONV_S1, c1, p->w[2], p->shift[2], 1, nullptr, c2, 0, wsc + 2 * kWScratch, B, 16, 16, D / 2, H / 2, W / 2, num_sms, st, 0, 0, true, c2s));
    RUN(run_layer(TC_CONV_S2, c2s ? c2s : c2, p->w[3], p->shift[3], 1, nullptr, c3, 0, wsc + 3 * kWScratch, B, 16, 32, D / 2, H / 2, W / 2, num_sms, st, 0, 0, true, nullptr, c2s != nullptr));
    RUN(run_layer(TC_CONV_S1, c3, p->w[4], p->shift[4], 1, nullptr, c4, 0, wsc + 4 * kWScratch, B, 32, 32, D / 4, H / 4, W / 4, num_sms, st, 0, 0, true, c4s));
    RUN(run_layer(TC_CONV_S2, c4s ? c4s : c4, p->w[5], p->shift[5], 1, nullptr, c5, 0, wsc + 5 * kWScratch, B, 32, 64, D / 4, H / 4, W / 4, num_sms, st, 0, 0, true, nullptr, c4s != nullptr));
    RUN(run_layer(TC_CONV_S1, c5, p->w[6], p->shift[6], 1, nullptr, c6, 0, wsc + 6 * kWScratch, B, 64, 64, D / 8, H / 8, W / 8, num_sms, st, 0, 0, true));
    RUN(run_layer(TC_CONVT, c6, p->w[7], p->shift[7], 1, c4, u7, 0, wsc + 7 * kWScratch, B, 64, 32, D / 8, H / 8, W / 8, num_sms, st, 0, 0, true));
    RUN(run_layer(TC_CONVT, u7, p->w[8], p->shift[8], 1, c2, u9, 0, wsc + 8 * kWScratch, B, 32, 16, D / 4, H / 4, W / 4, num_sms, st, 0, 0, true));
    RUN(run_layer(TC_CONVT, u9, p->w[9], p->shift[9], 1, c0, u11, 0, wsc + 9 * kWScratch, B, 16, 8, D / 2, H / 2, W / 2, num_sms, st, 0, 0, true));
    RUN(run_layer(TC_CONV_S1, u11, p->w[10], p->shift[10], 0, nullptr, logits, 1, wsc + 10 * kWScratch, B, 8, 1, D, H, W, num_sms, st, 0, 0, true));
#undef RUN
    return MVS_OK;
}

// Single-layer entry for tests: fp32 NCDHW in/out, converts around one tensor-core layer.
int tc_layer_ncdhw(int kind, const float *x, const float *w, const float *shift, int relu, const float *skip, float *y,
                   int B, int cin, int cout, int D, int H, int W, cudaStream_t st) {
    int dev = 0, num_sms = 148;
    MVS_CUDA(cudaGetDevice(&dev));
    MVS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t Nin = (size_t)D * H * W;
    int Do = D, Ho = H, Wo = W;
    if (kind == TC_CONV_S2) { Do = D / 2; Ho = H / 2; Wo = W / 2; }
    if (kind == TC_CONVT) { Do = 2 * D; Ho = 2 * H; Wo = 2 * W; }
    const size_t Nout = (size_t)Do * Ho * Wo;
    const int cout_pad = (cout + 7) / 8 * 8;
    const size_t in_b = align_up((size_t)B * cin * Nin * 2, 1024), out_b = align_up((size_t)B * cout_pad * Nout * 2, 1024);
    uint8_t *ws = nullptr;
    MVS_CUDA(cudaMallocAsync((void **)&ws, in_b + 2 * out_b + kWScratch, st));
    void *xin = ws, *yout = ws + in_b, *sk = ws + in_b + out_b, *wsc = ws + in_b + 2 * out_b;
    int rc = MVS_OK;
    ncdhw_to_cp8_kernel<<<dim3(cdiv(Nin, 256), B * cin / 8), 256, 0, st>>>(x, (uint4 *)xin, cin, Nin);
    if (skip) ncdhw_to_cp8_kernel<<<dim3(cdiv(Nout, 256), B * cout / 8), 256, 0, st>>>(skip, (uint4 *)sk, cout, Nout);
    const bool f32out = (cout == 1);
    rc = run_layer((TcKind)kind, xin, w, shift, relu, skip ? sk : nullptr, f32out ? (void *)y : yout, f32out, wsc, B, cin,
                   cout, D, H, W, num_sms, st);
    if (rc == MVS_OK && !f32out)
        cp8_to_ncdhw_kernel<<<dim3(cdiv(Nout, 256), B * cout / 8), 256, 0, st>>>((const uint4 *)yout, y, cout, Nout);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(ws, st);
    if (rc != MVS_OK) return rc;
    if (e != cudaSuccess) return set_error(MVS_ERR_CUDA, "tc_layer launch failed: %s", cudaGetErrorString(e));
    count_launches(3);
    return MVS_OK;
}

// ------------------------------------------------------------------------------------------------
// FeatureNet.forward (mvsnet.py:10-30) on the tensor cores, eval mode (BN folded by the caller), fp16 operands.
// Activations are fp16 "CP8 over images": [C/8][N][H][W][8] -- the N images play the role of the depth planes of
// the 3-D layers (no taps across them).  The two 5x5 stride-2 layers (conv2, conv5) are evaluated as 3x3
// stride-1 layers over the space-to-depth form of their input: out(yo,xo) = sum_{kh,kw} in(2yo+kh-2, 2xo+kw-2)
// w(kh,kw) with kh = 2 KH + py: the 3x3 window KH over half-resolution rows Y = yo-1..yo+1 of the four
// parity sub-images (py, px), i.e. a 3x3 conv with 4 Cin channels and the 5x5 weights scattered into 6x6
// (row/column 5 zero).  The layer before each of them writes that layout straight from its epilogue.
// ------------------------------------------------------------------------------------------------
// fp32 (or 8-bit, see below) NCHW [N][C][H][W] -> fp16 [ceil(C/8) (x4 if s2d)][N][Ho][Wo][8]; channels beyond C are zero.
// T = uint8_t: the image as decoded from disk; value / 255 in fp32 is what the reference's loader computes on the
// host before the upload (datasets/data_io.py: np.array(img, dtype=np.float32) / 255.), done here after a 4x smaller copy.
template <typename T>
__device__ __forceinline__ float load_pixel(const T *p) {
    if constexpr (sizeof(T) == 1) return __fdiv_rn((float)__ldg(p), 255.f);
    else return __ldg(p);
}
template <typename T>
__global__ void nchw_to_cp8n_f16_kernel(const T *__restrict__ in, uint4 *__restrict__ out, int N, int C, int H, int W,
                                        int s2d, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int Ho = s2d ? H / 2 : H, Wo = s2d ? W / 2 : W, cpc = (C + 7) / 8;
    const int x = (int)(i % Wo);
    long long r = i / Wo;
    const int y = (int)(r % Ho); r /= Ho;
    const int n = (int)(r % N);
    const int chunk = (int)(r / N);
    const int par = s2d ? chunk / cpc : 0, cc = s2d ? chunk % cpc : chunk;
    const int ys = s2d ? 2 * y + (par >> 1) : y, xs = s2d ? 2 * x + (par & 1) : x;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float f[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = cc * 8 + 2 * j + h;
            f[h] = (c < C) ? load_pixel(in + (((size_t)n * C + c) * H + ys) * W + xs) : 0.f;
        }
        w[j] = pack_f16x2(f[0], f[1]);
    }
    out[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// FeatureNet input: 3-channel images [N][3][H][W] (fp32, or uint8 / 255) -> fp16 [1][N][H][W][8] with channels 3..7 zero.
// Four consecutive pixels per thread: one 16-byte (fp32) or 4-byte (uint8) load per colour plane, 64 contiguous bytes out.
template <typename T>
__global__ void images_to_cp8_kernel(const T *__restrict__ in, uint4 *__restrict__ out, size_t HW, long long quads) {
    ptx::pdl_launch_dependents();  // the first FeatureNet layer may set up while this grid drains
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= quads) return;
    const size_t px = (size_t)i * 4;  // first pixel (n*HW + y*W + x); W % 4 == 0 keeps the four in one row
    const size_t n = px / HW, off = px - n * HW;
    float c[3][4];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const T *p = in + (n * 3 + ch) * HW + off;
        if constexpr (sizeof(T) == 1) {
            const uchar4 v = __ldg(reinterpret_cast<const uchar4 *>(p));
            c[ch][0] = __fdiv_rn((float)v.x, 255.f); c[ch][1] = __fdiv_rn((float)v.y, 255.f);
            c[ch][2] = __fdiv_rn((float)v.z, 255.f); c[ch][3] = __fdiv_rn((float)v.w, 255.f);
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
            c[ch][0] = v.x; c[ch][1] = v.y; c[ch][2] = v.z; c[ch][3] = v.w;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(out + px + k, make_uint4(pack_f16x2(c[0][k], c[1][k]), pack_f16x2(c[2][k], 0.f), 0u, 0u));
}

// fp16 [C/8][N][H][W][8] -> fp32 NCHW (tests)
__global__ void cp8n_f16_to_nchw_kernel(const uint4 *__restrict__ in, float *__restrict__ out, int N, int C, int H, int W,
                                        long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W);
    long long r = i / W;
    const int y = (int)(r % H); r /= H;
    const int n = (int)(r % N);
    const int cc = (int)(r / N);
    const uint4 v = __ldg(in + i);
    const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        out[(((size_t)n * C + cc * 8 + 2 * j) * H + y) * W + x] = f.x;
        out[(((size_t)n * C + cc * 8 + 2 * j + 1) * H + y) * W + x] = f.y;
    }
}

// Effective 3x3 weights [Cout][Cin_eff][3][3] of one FeatureNet layer from its native [Cout][Cin][k][k]:
// k = 3: channels padded with zeros up to Cin_eff;  k = 5 (stride 2): the space-to-depth form, Cin_eff = 4 Cin,
// effective channel = (py*2+px)*Cin + ci, tap (KH,KW) <- (kh,kw) = (2KH+py, 2KW+px) when both are < 5.
__global__ void featurenet_effective_weights_kernel(const float *__restrict__ w, float *__restrict__ out, int cout, int cin,
                                                    int cin_eff, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cout * cin_eff * 9) return;
    const int KW = i % 3, KH = (i / 3) % 3, ce = (i / 9) % cin_eff, co = i / (9 * cin_eff);
    float v = 0.f;
    if (k == 3) {
        if (ce < cin) v = w[((size_t)(co * cin + ce) * 3 + KH) * 3 + KW];
    } else {
        const int par = ce / cin, ci = ce % cin, kh = 2 * KH + (par >> 1), kw = 2 * KW + (par & 1);
        if (kh < 5 && kw < 5) v = w[((size_t)(co * cin + ci) * 5 + kh) * 5 + kw];
    }
    out[i] = v;
}

// layer table: native Cin, Cout, kernel size (stride 2 when 5), effective Cin of the 3x3 form
static const int kFeatCin[MVS_FEATURENET_LAYERS] = {3, 8, 8, 16, 16, 16, 32, 32};
static const int kFeatCout[MVS_FEATURENET_LAYERS] = {8, 8, 16, 16, 16, 32, 32, 32};
static const int kFeatK[MVS_FEATURENET_LAYERS] = {3, 3, 5, 3, 3, 5, 3, 3};
static const int kFeatCinEff[MVS_FEATURENET_LAYERS] = {8, 8, 32, 16, 16, 64, 32, 32};

size_t featurenet_tc_workspace_bytes(int N, int H, int W) {
    // per input pixel and image: in 16 | a0 16 | a1 (s2d) 16 | a2 8 | a3 8 | a4 (s2d) 8 | a5 4 | a6 4 bytes
    return align_up((size_t)N * H * W * 80, 1024) + 8 * 1024 + MVS_FEATURENET_LAYERS * (kWScratch + 128 * 1024) + 16 * 1024;
}

// imgs fp32 [N][3][H][W] -> fea fp16 RCP8 [N][H/4][4][W/4][8]
int featurenet_tc(const void *imgs, int imgs_u8, const mvs_featurenet_params *p, void *fea, void *workspace, int N, int H,
                  int W, cudaStream_t st) {
    std::shared_lock<std::shared_mutex> lease(g_wcache_life);  // cached weight pointers stay valid until the last launch
    int dev = 0, num_sms = 148;
    MVS_CUDA(cudaGetDevice(&dev));
    MVS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t px = (size_t)N * H * W;
    uint8_t *ws = (uint8_t *)workspace;
    auto take = [&](size_t bytes) { uint8_t *r = ws; ws += align_up(bytes, 1024); return r; };
    void *in0 = take(px * 16), *a0 = take(px * 16), *a1 = take(px * 16), *a2 = take(px * 8), *a3 = take(px * 8),
         *a4 = take(px * 8), *a5 = take(px * 4), *a6 = take(px * 4);
    uint8_t *wsc = take(MVS_FEATURENET_LAYERS * (kWScratch + 128 * 1024));
    {
        const long long total = (long long)px;
        const long long quads = total / 4;  // W % 4 == 0
        const bool aligned = ((uintptr_t)imgs & 15) == 0;
        // the vectorised kernels read 4 pixels per load: uchar4 needs a 4-byte aligned base (a uint8 view with an odd
        // storage offset is legal for the caller), float4 a 16-byte aligned one; otherwise the scalar kernel
        if (imgs_u8 && ((uintptr_t)imgs & 3) != 0) nchw_to_cp8n_f16_kernel<uint8_t><<<cdiv(total, 256), 256, 0, st>>>((const uint8_t *)imgs, (uint4 *)in0, N, 3, H, W, 0, total);
        else if (imgs_u8) images_to_cp8_kernel<uint8_t><<<cdiv(quads, 256), 256, 0, st>>>((const uint8_t *)imgs, (uint4 *)in0, (size_t)H * W, quads);
        else if (aligned) images_to_cp8_kernel<float><<<cdiv(quads, 256), 256, 0, st>>>((const float *)imgs, (uint4 *)in0, (size_t)H * W, quads);
        else nchw_to_cp8n_f16_kernel<float><<<cdiv(total, 256), 256, 0, st>>>((const float *)imgs, (uint4 *)in0, N, 3, H, W, 0, total);
        MVS_LAUNCH_CHECK(1);
    }
    const void *ins[MVS_FEATURENET_LAYERS] = {in0, a0, a1, a2, a3, a4, a5, a6};
    void *outs[MVS_FEATURENET_LAYERS] = {a0, a1, a2, a3, a4, a5, a6, fea};
    const int scale[MVS_FEATURENET_LAYERS] = {1, 1, 2, 2, 2, 4, 4, 4};   // resolution divisor of each layer's input
    const int out_mode[MVS_FEATURENET_LAYERS] = {0, 1, 0, 0, 1, 0, 0, 2};  // last layer: RCP8 for the warp kernel
    for (int l = 0; l < MVS_FEATURENET_LAYERS; ++l) {
        void *wpk = wsc;  // unused since the packed weights are cached; kept for the signature
        const int nw = kFeatCout[l] * kFeatCinEff[l] * 9;
        void *weff_v = nullptr;
        const WCacheKey key{dev, p->w[l], 0xFEA70000ull + (uint64_t)l};
        if (int rc = wcache_get(key, (size_t)nw * sizeof(float), st, &weff_v, [&](void *dst) -> int {
                featurenet_effective_weights_kernel<<<cdiv(nw, 256), 256, 0, st>>>(p->w[l], (float *)dst, kFeatCout[l], kFeatCin[l],
                                                                                  kFeatCinEff[l], kFeatK[l]);
                MVS_LAUNCH_CHECK(1);
                return MVS_OK;
            }))
            return rc;
        const float *weff = (const float *)weff_v;
        if (int rc = run_layer(TC_CONV2D, ins[l], weff, p->shift[l], l != MVS_FEATURENET_LAYERS - 1, nullptr, outs[l], 0, wpk,
                               1, kFeatCinEff[l], kFeatCout[l], N, H / scale[l], W / scale[l], num_sms, st, 1, out_mode[l], true))
            return rc;
    }
    return MVS_OK;
}

// Single 2-D layer for tests: fp32 NCHW in/out.  ksize 3 (stride 1) or 5 (stride 2, through the space-to-depth form).
int tc_conv2d_nchw(const float *x, const float *w, const float *shift, int relu, float *y, int N, int cin, int cout, int H,
                   int W, int ksize, int s2d_out, cudaStream_t st) {
    int dev = 0, num_sms = 148;
    MVS_CUDA(cudaGetDevice(&dev));
    MVS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    const int s2d_in = (ksize == 5);
    const int cin8 = (cin + 7) / 8 * 8, cin_eff = s2d_in ? 4 * cin8 : cin8;
    const int Hi = s2d_in ? H / 2 : H, Wi = s2d_in ? W / 2 : W;
    const size_t in_b = align_up((size_t)N * Hi * Wi * cin_eff * 2, 1024), out_b = align_up((size_t)N * Hi * Wi * cout * 2, 1024);
    uint8_t *ws = nullptr;
    MVS_CUDA(cudaMallocAsync((void **)&ws, in_b + out_b + kWScratch + 128 * 1024, st));
    void *xin = ws, *yout = ws + in_b;
    float *weff = (float *)(ws + in_b + out_b);
    void *wpk = (uint8_t *)weff + 128 * 1024;
    const long long tin = (long long)N * Hi * Wi * (cin_eff / 8);
    nchw_to_cp8n_f16_kernel<float><<<cdiv(tin, 256), 256, 0, st>>>(x, (uint4 *)xin, N, cin, H, W, s2d_in, tin);
    // with cin not a multiple of 8 the s2d channel order is parity * cin8 + ci: build the weights on the padded count
    featurenet_effective_weights_kernel<<<cdiv(cout * cin_eff * 9, 256), 256, 0, st>>>(w, weff, cout, cin, cin_eff, ksize);
    int rc = (s2d_in && cin != cin8) ? set_error(MVS_ERR_UNSUPPORTED, "5x5 stride-2 layer needs Cin %% 8 == 0") : MVS_OK;
    if (rc == MVS_OK)
        rc = run_layer(TC_CONV2D, xin, weff, shift, relu, nullptr, yout, 0, wpk, 1, cin_eff, cout, N, Hi, Wi, num_sms, st, 1, s2d_out);
    if (rc == MVS_OK) {
        const int Co = s2d_out ? 4 * cout : cout, Ho = s2d_out ? Hi / 2 : Hi, Wo = s2d_out ? Wi / 2 : Wi;
        const long long tout = (long long)N * Ho * Wo * (Co / 8);
        cp8n_f16_to_nchw_kernel<<<cdiv(tout, 256), 256, 0, st>>>((const uint4 *)yout, y, N, Co, Ho, Wo, tout);
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(ws, st);
    if (rc != MVS_OK) return rc;
    if (e != cudaSuccess) return set_error(MVS_ERR_CUDA, "tc_conv2d launch failed: %s", cudaGetErrorString(e));
    count_launches(3);
    return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_conv3d_bn_relu_tc(const float *x, const float *w, const float *shift, int relu, float *y, int B,
                                     int Cin, int Cout, int D, int H, int W, int stride, void *stream) {
    MVS_REQUIRE(x && w && shift && y, "null pointer argument");
    MVS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "bad shape");
    MVS_REQUIRE(stride == 1 || stride == 2, "conv3d: stride must be 1 or 2, got %d", stride);
    MVS_REQUIRE(Cin % 8 == 0 && (Cout % 8 == 0 || Cout == 1), "tensor-core conv3d needs Cin %% 8 == 0 and Cout %% 8 == 0 (or 1)");
    MVS_REQUIRE(stride == 1 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "stride-2 tensor-core conv3d needs even D, H, W");
    return tc_layer_ncdhw(stride == 1 ? TC_CONV_S1 : TC_CONV_S2, x, w, shift, relu, nullptr, y, B, Cin, Cout, D, H, W,
                          (cudaStream_t)stream);
}

extern "C" int mvs_conv_transpose3d_bn_relu_tc(const float *x, const float *w, const float *shift, int relu,
                                               const float *skip, float *y, int B, int Cin, int Cout, int D, int H,
                                               int W, void *stream) {
    MVS_REQUIRE(x && w && shift && y, "null pointer argument");
    MVS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "bad shape");
    MVS_REQUIRE(Cin % 16 == 0 && Cout % 8 == 0, "tensor-core conv_transpose3d needs Cin %% 16 == 0 and Cout %% 8 == 0");
    return tc_layer_ncdhw(TC_CONVT, x, w, shift, relu, skip, y, B, Cin, Cout, D, H, W, (cudaStream_t)stream);
}

// Diagnostics: device buffer of [grid][12] int64 cycle counters filled by the next tensor-core layer launches
// (producer wait, MMA wait-full / wait-tmem / issue / total / steps, epilogue wait / work).  nullptr disables.
extern "C" int mvs_tc_set_debug_buffer(void *buf) {
    g_tc_dbg = (long long *)buf;
    return MVS_OK;
}

// Diagnostics: the tile / ring / grid configuration the planner picks for one layer (no GPU needed).
extern "C" int mvs_tc_plan_describe(int kind, int B, int Cin, int Cout, int D, int H, int W, int num_sms, char *buf,
                                    int buflen) {
    static thread_local TcPlan pl;
    MVS_REQUIRE(buf && buflen > 0 && kind >= 0 && kind <= 2, "bad argument");
    if (int rc = make_plan(pl, (TcKind)kind, B, Cin, Cout, D, H, W, nullptr, num_sms, false)) return rc;
    const TcLayer &L = pl.L;
    snprintf(buf, buflen,
             "kind=%d cin=%d cout=%d tile=%dx%d (P=%d) MT=%d nslot=%d need=%d adv=%d nsub=%d slot=%dB w=%dB smem=%zuB "
             "npad=%d groups=%d nacc=%d ops=%d tiles=%dx%d zsegs=%d(len %d) items=%d grid=%d useful=%.3f",
             kind, Cin, Cout, L.TXB, L.TY, L.P, L.MT, L.nslot, L.need, L.adv, L.nsub, L.slot_bytes, L.wbytes_group,
             pl.smem_bytes, L.npad, L.ngroups, L.nacc, L.nops, L.tiles_x, L.tiles_y, L.zsegs, L.zseg_len, L.n_items,
             pl.grid, (double)L.Wt * L.Ht / ((double)L.tiles_x * L.tiles_y * L.MT * 128));
    return MVS_OK;
}

// ---- FeatureNet on the tensor cores (mvsnet.py:10-30, eval mode) --------------------------------
extern "C" size_t mvs_featurenet_tc_workspace_bytes(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0 || (H % 4) || (W % 4)) return 0;
    return featurenet_tc_workspace_bytes(N, H, W);
}

extern "C" int mvs_featurenet_tc_fwd(const float *imgs, const mvs_featurenet_params *params, void *fea_rcp8_f16,
                                     void *workspace, int N, int H, int W, void *stream) {
    MVS_REQUIRE(imgs && params && fea_rcp8_f16 && workspace, "null pointer argument");
    MVS_REQUIRE(N > 0 && H >= 4 && W >= 4 && H % 4 == 0 && W % 4 == 0,
                "FeatureNet needs H, W divisible by 4 (two stride-2 stages), got %dx%d", H, W);
    for (int i = 0; i < MVS_FEATURENET_LAYERS; ++i)
        MVS_REQUIRE(params->w[i] && params->shift[i], "featurenet params: layer %d has a null pointer", i);
    return featurenet_tc(imgs, 0, params, fea_rcp8_f16, workspace, N, H, W, (cudaStream_t)stream);
}

extern "C" int mvs_featurenet_tc_fwd_u8(const uint8_t *imgs_u8, const mvs_featurenet_params *params, void *fea_rcp8_f16,
                                        void *workspace, int N, int H, int W, void *stream) {
    MVS_REQUIRE(imgs_u8 && params && fea_rcp8_f16 && workspace, "null pointer argument");
    MVS_REQUIRE(N > 0 && H >= 4 && W >= 4 && H % 4 == 0 && W % 4 == 0,
                "FeatureNet needs H, W divisible by 4 (two stride-2 stages), got %dx%d", H, W);
    for (int i = 0; i < MVS_FEATURENET_LAYERS; ++i)
        MVS_REQUIRE(params->w[i] && params->shift[i], "featurenet params: layer %d has a null pointer", i);
    return featurenet_tc(imgs_u8, 1, params, fea_rcp8_f16, workspace, N, H, W, (cudaStream_t)stream);
}

extern "C" int mvs_conv2d_bn_relu_tc(const float *x, const float *w, const float *shift, int relu, float *y, int N, int Cin,
                                     int Cout, int H, int W, int ksize, int stride, int s2d_out, void *stream) {
    MVS_REQUIRE(x && w && shift && y, "null pointer argument");
    MVS_REQUIRE(N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "bad shape");
    MVS_REQUIRE((ksize == 3 && stride == 1) || (ksize == 5 && stride == 2), "conv2d: 3x3 stride 1 or 5x5 stride 2 only");
    MVS_REQUIRE(Cout % 8 == 0 && Cout <= 32, "tensor-core conv2d needs Cout in {8, 16, 24, 32}");
    MVS_REQUIRE(ksize == 3 || (H % 2 == 0 && W % 2 == 0), "stride-2 conv2d needs even H, W");
    return tc_conv2d_nchw(x, w, shift, relu, y, N, Cin, Cout, H, W, ksize, s2d_out, (cudaStream_t)stream);
}

/* Drops every cached packed-weight buffer (see the cache comment above run_layer). */
extern "C" int mvs_weight_cache_clear(void) { return weight_cache_clear(); }
