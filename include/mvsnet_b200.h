/*
 * mvsnet_b200.h -- C ABI of the B200-native MVSNet depth-inference hot path.
 *
 * One shared library (libmvsnet_b200.so, hand-written CUDA for sm_100a) exports exactly these
 * symbols.  Signatures use plain pointers and sizes only -- no torch types -- so the same library
 * is bound from Python (ctypes, see scene_3dreconstruction_mvsnet_b200/_lib.py and INTEGRATION.md),
 * C or C++.  Each entry point names the reference interface it replaces (paths relative to the
 * reference repository olivier-2018/scene_3Dreconstruction_MVSNet).
 *
 * Conventions
 *   - All tensors are dense fp32 in the reference's layouts: feature maps NCHW, volumes NCDHW,
 *     projection matrices row-major 4x4 with K[R|t] in rows 0-2 (datasets/dataloader_eval.py:158-159).
 *   - Pointers named *_host are host memory; every other data pointer is DEVICE memory on the
 *     current device.  `stream` is a cudaStream_t (NULL = legacy default stream).  Calls are
 *     asynchronous with respect to the host unless the name ends in _host.
 *   - Inputs are borrowed and never written.  Outputs/workspaces are caller-allocated.
 *   - Every function returns MVS_OK (0) or a negative mvs_status; mvs_last_error() gives the
 *     message for the calling thread.  Nothing aborts the process.  There is no CPU fallback.
 *   - Re-entrant: safe to call from several host threads on several devices (nn.DataParallel calls the
 *     reference's forward from one thread per device, train.py:125).  Global state: an atomic launch counter and
 *     the mutex-protected packed-weight / launch-plan cache of the tensor-core layers (mvs_weight_cache_clear).
 */
#ifndef MVSNET_B200_H
#define MVSNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVSNET_B200_ABI_VERSION 2

typedef enum {
    MVS_OK = 0,
    MVS_ERR_INVALID_ARG = -1, /* shape / pointer / alignment contract violated (the reference would raise) */
    MVS_ERR_CUDA = -2,        /* a CUDA runtime call or kernel launch failed */
    MVS_ERR_UNSUPPORTED = -3  /* valid for the reference but outside this build's scope */
} mvs_status;

/* Precision of the CostRegNet contraction.  FP32 = CUDA-core fp32 FMA, matches the reference to
 * fp32 rounding.  BF16 (the name is historical) = tcgen05 tensor-core implicit GEMM with 16-bit operands -- fp16
 * since ABI 2, activations saturating at 65504 -- and fp32 accumulation (looser, separately stated tolerance). */
typedef enum { MVS_PRECISION_FP32 = 0, MVS_PRECISION_BF16 = 1 } mvs_precision;

int mvs_abi_version(void);
const char *mvs_last_error(void);
/* Number of kernels this library has launched so far in this process (all threads). */
uint64_t mvs_launch_count(void);
/* Compiled-for architecture string, e.g. "sm_100a". */
const char *mvs_arch(void);

/* ---- (a2) homo_warping(src_fea, src_proj, ref_proj, depth_values)      models/module.py:96-139
 * src_fea [B,C,H,W], src_proj/ref_proj [B,4,4], depth_values [B,D]  ->  out [B,C,D,H,W].
 * Bilinear, zero padding, the reference's align_corners mismatch reproduced (module.py:130-136). */
int mvs_homo_warping(const float *src_fea, const float *src_proj, const float *ref_proj, const float *depth_values,
                     float *out, int B, int C, int D, int H, int W, void *stream);

/* Backward of homo_warping w.r.t. src_fea (the grid carries no gradient, module.py:106).
 * grad_out [B,C,D,H,W] -> grad_src [B,C,H,W] (overwritten). */
int mvs_homo_warping_bwd(const float *grad_out, const float *src_proj, const float *ref_proj,
                         const float *depth_values, float *grad_src, int B, int C, int D, int H, int W,
                         void *stream);

/* ---- (a2+a3) fused plane-sweep warp + variance cost volume             models/mvsnet.py:145-177
 * fea [B,V,C,H,W] (view 0 = reference view), proj [B,V,4,4], depth_values [B,D]
 *   -> var [B,C,D,H,W] = sum(x^2)/V - (sum(x)/V)^2 over the V views; per-view warped volumes are
 * never written to memory.  C must be 32 (FeatureNet's width, mvsnet.py:24).
 * workspace: mvs_warp_variance_workspace_bytes() bytes of device memory (channels-last copy of the
 * source-view features + the composed homographies). */
size_t mvs_warp_variance_workspace_bytes(int B, int V, int C, int H, int W);
int mvs_warp_variance_fwd(const float *fea, const float *proj, const float *depth_values, float *var,
                          void *workspace, int B, int V, int C, int D, int H, int W, void *stream);

/* ---- (a8) backward of the fused op (autograd through mvsnet.py:167-177 + grid_sample)
 * grad_var [B,C,D,H,W] -> grad_fea [B,V,C,H,W] (overwritten; view 0 = reference view).
 * workspace: mvs_warp_variance_bwd_workspace_bytes() bytes. */
size_t mvs_warp_variance_bwd_workspace_bytes(int B, int V, int C, int H, int W);
int mvs_warp_variance_bwd(const float *grad_var, const float *fea, const float *proj, const float *depth_values,
                          float *grad_fea, void *workspace, int B, int V, int C, int D, int H, int W,
                          void *stream);

/* ---- (a4) CostRegNet building blocks                    models/module.py:26-33, models/mvsnet.py:33-73
 * Conv3d k=3 pad=1 stride 1|2 with eval-mode BatchNorm folded by the caller:
 *   y = act( conv(x, w) + shift ),  w [Cout,Cin,3,3,3] already multiplied by gamma/sqrt(var+eps),
 *   shift [Cout] = beta - mean*gamma/sqrt(var+eps) (or the conv bias for the final prob layer). */
int mvs_conv3d_bn_relu(const float *x, const float *w, const float *shift, int relu, float *y, int B, int Cin,
                       int Cout, int D, int H, int W, int stride, void *stream);
/* ConvTranspose3d k=3 stride=2 pad=1 output_padding=1 (mvsnet.py:46-59), BN folded the same way,
 * w [Cin,Cout,3,3,3]; y = skip + act(convT(x,w) + shift) with skip optional (mvsnet.py:69-71).
 * x [B,Cin,D,H,W] -> y [B,Cout,2D,2H,2W]. */
int mvs_conv_transpose3d_bn_relu(const float *x, const float *w, const float *shift, int relu, const float *skip,
                                 float *y, int B, int Cin, int Cout, int D, int H, int W, void *stream);

/* Tensor-core variants of the two building blocks (tcgen05 implicit GEMM, fp16 operands, fp32
 * accumulate; same fp32 NCDHW interface, converted internally).  Cin % 8 == 0, Cout % 8 == 0 or 1. */
int mvs_conv3d_bn_relu_tc(const float *x, const float *w, const float *shift, int relu, float *y, int B, int Cin,
                          int Cout, int D, int H, int W, int stride, void *stream);
int mvs_conv_transpose3d_bn_relu_tc(const float *x, const float *w, const float *shift, int relu, const float *skip,
                                    float *y, int B, int Cin, int Cout, int D, int H, int W, void *stream);

/* The tensor-core layers pack their weights (fp32 -> 16-bit MMA operand blocks) once per weight POINTER and layer
 * configuration and reuse the packed copy in later calls.  Call this after changing weights in place or freeing and
 * re-creating them (nn.Module.load_state_dict, an optimizer step, re-folding BatchNorm). */
int mvs_weight_cache_clear(void);

/* Diagnostics (no GPU needed): describes the tile/ring/grid plan of one tensor-core layer.
 * kind: 0 = conv stride 1, 1 = conv stride 2, 2 = transposed conv. */
int mvs_tc_set_debug_buffer(void *device_buf_or_null); /* [grid][12] int64 cycle counters of the next tc launches */
int mvs_tc_plan_describe(int kind, int B, int Cin, int Cout, int D, int H, int W, int num_sms, char *buf, int buflen);

/* Whole CostRegNet.forward (mvsnet.py:64-73), eval mode.  Layer order in the arrays:
 * conv0..conv6, conv7, conv9, conv11, prob.  volume [B,32,D,H,W] -> logits [B,D,H,W]. */
#define MVS_COSTREG_LAYERS 11
typedef struct {
    const float *w[MVS_COSTREG_LAYERS];     /* folded weights, layouts as above (device) */
    const float *shift[MVS_COSTREG_LAYERS]; /* folded shifts (device) */
} mvs_costreg_params;
size_t mvs_costreg_workspace_bytes(int B, int D, int H, int W, int precision);
int mvs_costreg_fwd(const float *volume, const mvs_costreg_params *params, float *logits, void *workspace, int B,
                    int D, int H, int W, int precision, void *stream);

/* tensor-core precision mode, fused layout: the warp+variance kernel writes the cost volume directly as fp16
 * "CP8" [B][32/8][D][H][W][8] (mvs_volume_cp8_bytes bytes), which mvs_costreg_fwd_cp8 consumes on the
 * tensor cores -- the fp32 volume and its conversion pass never exist.  Workspaces as for the fp32 calls
 * (mvs_warp_variance_workspace_bytes / mvs_costreg_workspace_bytes with MVS_PRECISION_BF16). */
size_t mvs_volume_cp8_bytes(int B, int D, int H, int W);
int mvs_warp_variance_fwd_cp8(const float *fea, const float *proj, const float *depth_values, void *vol_cp8,
                              void *workspace, int B, int V, int C, int D, int H, int W, void *stream);
/* Same, but the features arrive as ONE fp16 channels-last tensor [B][V][H][W][32] (what a half-precision FeatureNet
 * emits): no layout pre-pass; view 0 is the reference view. */
int mvs_warp_variance_fwd_cp8_f16(const void *fea16_nhwc, const float *proj, const float *depth_values, void *vol_cp8,
                                  void *workspace, int B, int V, int C, int D, int H, int W, void *stream);
int mvs_costreg_fwd_cp8(const void *vol_cp8, const mvs_costreg_params *params, float *logits, void *workspace, int B,
                        int D, int H, int W, void *stream);

/* ---- FeatureNet.forward (models/mvsnet.py:10-30), eval mode, on the tensor cores (fp16 operands, fp32 accumulate).
 * the stage next to the hot path.  imgs [N,3,H,W] fp32 (N = B*V images) ->
 * fea_rcp8_f16: fp16 row-chunk-planar [N][H/4][32/8][W/4][8], the layout mvs_warp_variance_fwd_cp8_feat samples
 * through its TMA windows (no conversion pass in between).
 * params: the 8 layers conv0, conv1, conv2, conv3, conv4, conv5, conv6, feature in their NATIVE shapes
 * ([Cout][Cin][k][k], k = 5 for conv2 / conv5), eval-mode BN already folded: w scaled, shift = folded bias
 * (the last layer's shift is its conv bias).  H, W divisible by 4. */
#define MVS_FEATURENET_LAYERS 8
typedef struct {
    const float *w[MVS_FEATURENET_LAYERS];
    const float *shift[MVS_FEATURENET_LAYERS];
} mvs_featurenet_params;
size_t mvs_featurenet_tc_workspace_bytes(int N, int H, int W);
int mvs_featurenet_tc_fwd(const float *imgs, const mvs_featurenet_params *params, void *fea_rcp8_f16, void *workspace,
                          int N, int H, int W, void *stream);
/* Same, images as 8-bit [N,3,H,W] (as decoded from disk): value/255 in fp32 -- what the reference's loader computes on
 * the host (datasets/data_io.py) -- happens on the device after a 4x smaller host->device copy. */
int mvs_featurenet_tc_fwd_u8(const uint8_t *imgs_u8, const mvs_featurenet_params *params, void *fea_rcp8_f16,
                             void *workspace, int N, int H, int W, void *stream);
/* One ConvBnReLU (models/module.py:8-15) on the same kernel, fp32 NCHW in/out (tests, diagnostics):
 * ksize 3 / stride 1 / pad 1, or ksize 5 / stride 2 / pad 2.  s2d_out = 1 returns the space-to-depth form
 * [N, 4*Cout, H'/2, W'/2] (channel = (y&1)*2+(x&1) major) that a following stride-2 layer consumes. */
int mvs_conv2d_bn_relu_tc(const float *x, const float *w, const float *shift, int relu, float *y, int N, int Cin, int Cout,
                          int H, int W, int ksize, int stride, int s2d_out, void *stream);
/* ---- FeatureNet.forward (models/mvsnet.py:10-30), eval mode, at the reference's precision: fp32 FMA on the CUDA cores
 * (TMA-staged halo tiles, BN folded, bias / ReLU in the epilogue).  imgs [N,3,H,W] fp32 -> fea [N,32,H/4,W/4] fp32 NCHW,
 * what mvs_warp_variance_fwd takes.  Same parameter struct as mvs_featurenet_tc_fwd.  H % 4 == 0 and W % 16 == 0 (every
 * layer's row pitch must be a multiple of 16 bytes for the tensor maps); mvs_featurenet_workspace_bytes() returns 0
 * otherwise.  workspace: mvs_featurenet_workspace_bytes() bytes. */
size_t mvs_featurenet_workspace_bytes(int N, int H, int W);
int mvs_featurenet_fwd(const float *imgs, const mvs_featurenet_params *params, float *fea, void *workspace, int N, int H,
                       int W, void *stream);
/* One ConvBnReLU (models/module.py:8-15) of that path: x [N,Cin,H,W] -> y [N,Cout,H',W'] fp32 NCHW, ksize 3 / stride 1 /
 * pad 1 or ksize 5 / stride 2 / pad 2, BN folded into w [Cout][Cin][k][k] and shift.  W % 4 == 0, x 16-byte aligned. */
int mvs_conv2d_bn_relu(const float *x, const float *w, const float *shift, int relu, float *y, int N, int Cin, int Cout, int H,
                       int W, int ksize, int stride, void *stream);
/* Fused warp+variance on features in the layout mvs_featurenet_tc_fwd produces: fea [B*V][H][4][W][8] fp16 with
 * image index n = b*V + v (view 0 = reference view).  workspace: mvs_warp_variance_workspace_bytes().
 * fp16 texels, packed-half interpolation and packed-half sums of the deviations from the reference view; the fp16
 * volume saturates at 65504 (|feature| up to ~250).  |var - ref| <= 2^-6 |ref| + 8e-3 on N(0,1) features. */
int mvs_warp_variance_fwd_cp8_feat(const void *fea_rcp8_f16, const float *proj, const float *depth_values, void *vol_cp8,
                                   void *workspace, int B, int V, int C, int D, int H, int W, void *stream);

/* Same, features taken from a POOL of images [n_pool][H][4][W][8] fp16: view v of batch element b is image
 * view_ids_host[b*V + v] (HOST array of B*V ints, B*V <= 32; read during the call).  This is the scan-level form of
 * the reference's eval loop (eval.py:326-360): there every reference view re-loads and re-extracts its source images
 * (datasets/dataloader_eval.py:101-176 -> mvsnet.py:125); with a pool each image of a scan passes FeatureNet once and
 * serves every reference view whose pair list names it. */
int mvs_warp_variance_fwd_cp8_pool(const void *pool_rcp8_f16, int n_pool, const int *view_ids_host, const float *proj,
                                   const float *depth_values, void *vol_cp8, void *workspace, int B, int V, int C, int D,
                                   int H, int W, void *stream);

/* ---- (a5-a7) softmax over depth + depth expectation + 4-plane photometric confidence
 *                                                        models/mvsnet.py:192-193,204,214-218
 * logits [B,D,H,W], depth_values [B,D] -> depth [B,H,W], conf [B,H,W]; prob [B,D,H,W] optional. */
int mvs_softmax_depth_conf(const float *logits, const float *depth_values, float *depth, float *conf, float *prob,
                           int B, int D, int H, int W, void *stream);

/* ---- (a6) depth_regression(p, depth_values)                           models/module.py:144-147
 * p [B,D,H,W]; depth_values [B,D] (dv_batch_stride = D) or a shared [D] vector (dv_batch_stride = 0,
 * the call at mvsnet.py:217) -> out [B,H,W]. */
int mvs_depth_regression(const float *p, const float *depth_values, int dv_batch_stride, float *out, int B, int D,
                         int H, int W, void *stream);

/* ---- Geometric-consistency filter of the depth maps, per reference view     eval.py:508-585 (reproject_with_depth,
 * check_geometric_consistency) and eval.py:660-703 (filter_depth: masks and averaged depth).  SURVEY.md 8(f) rank 3.
 * ref_depth, confidence [H,W] fp32 and src_depths [S,H,W] fp32 are DEVICE pointers (confidence may be NULL: photo mask
 * all true); the camera matrices are HOST pointers, row-major double: intrinsics 3x3, extrinsics 4x4 (world -> camera),
 * as read by the reference's read_camera_parameters.  Outputs (device): depth_avg [H,W] float64 (the reference's
 * depth_est_averaged is float64), geo_mask_sum [H,W] int32, photo/geo/final masks [H,W] uint8; optional (NULL to skip):
 * reprojected [S,H,W] fp32 (0 where inconsistent), src_masks [S,H,W] uint8, xy_src [S,2,H,W] fp32 (x2d_src, y2d_src).
 * Sampling reproduces cv2.remap(INTER_LINEAR): 1/32-pixel coordinate quantisation, constant border 0. */
int mvs_filter_depth(const float *ref_depth, const float *confidence, const double *ref_K_host, const double *ref_E_host,
                     const float *src_depths, const double *src_K_host, const double *src_E_host, int S, int H, int W,
                     double condmask_pixel, double condmask_depth, int geomask, double photomask, double *depth_avg,
                     int32_t *geo_mask_sum, uint8_t *photo_mask, uint8_t *geo_mask, uint8_t *final_mask, float *reprojected,
                     uint8_t *src_masks, float *xy_src, void *stream);

/* ---- Host-buffer entry point: features to depth map, everything this library owns in one call.
 * Copies fea/proj/depth_values host->device, runs warp+variance -> CostRegNet -> softmax/depth/conf,
 * copies depth/conf device->host and synchronises.  params_host holds HOST pointers to the folded
 * weights (element counts implied by the layer table).  Allocates and frees its own device memory. */
int mvs_depth_from_features_host(const float *fea_host, const float *proj_host, const float *depth_values_host,
                                 const mvs_costreg_params *params_host, float *depth_host, float *conf_host, int B,
                                 int V, int D, int H, int W, int precision, int device);

#ifdef __cplusplus
}
#endif
#endif /* MVSNET_B200_H */
