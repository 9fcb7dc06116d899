"""Torch-tensor front end of the C ABI: device memory, streams and autograd plumbing only.

Every function here takes CUDA fp32 tensors, hands raw device pointers and the current CUDA
stream to libmvsnet_b200.so, and returns freshly allocated tensors.  Nothing is computed in
PyTorch on the forward path; non-CUDA inputs raise (there is no CPU fallback).
"""
import ctypes

import torch

from . import _lib


def _prep(t, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: the B200 path has no CPU fallback (move inputs to a CUDA device)"
                           % (name, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32, got %s" % (name, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError("%s must have %d dims, got shape %s" % (name, ndim, tuple(t.shape)))
    return t.detach().contiguous()


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


_WS_CACHE = {}
KERNEL_EVENTS = None  # set to a list to collect (name, start event, end event) around the fused warp kernels (bench.py)


class _timed:
    """Records CUDA events around a native call when KERNEL_EVENTS is a list (kernel time inside autograd graphs)."""

    def __init__(self, name, device):
        self.name, self.device = name, device

    def __enter__(self):
        if KERNEL_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.device))

    def __exit__(self, *exc):
        if KERNEL_EVENTS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(self.device))
            KERNEL_EVENTS.append((self.name, self.e0, e1))
        return False


def _ws(nbytes, device, tag=None):
    """Scratch device memory.  With a tag the buffer is kept and reused by later calls on the same device and
    stream (stream-ordered reuse is safe; it spares the caching allocator multi-GB alloc/free churn per step)."""
    nbytes = max(int(nbytes), 16)
    if tag is None:
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (tag, torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        _WS_CACHE[key] = buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return buf


def weights_changed():
    """Tell the native library that weight tensors were replaced or modified: its packed-weight cache is keyed by
    pointer (mvs_weight_cache_clear in include/mvsnet_b200.h).  Called by the models whenever BatchNorm is re-folded."""
    import os
    if os.path.exists(_lib.LIB_PATH):
        _lib.load().mvs_weight_cache_clear()


def compose_like_reference(proj):
    """proj [B,V,4,4] -> [B,V,4,4] whose view 0 is the identity and whose view v is `src_proj @ inverse(ref_proj)`
    computed by the reference's own torch calls (models/module.py:107, fp32, on the tensors' device).  The library then
    composes H_v * inverse(I) = H_v exactly, so the strict-fp32 path samples at the reference's homographies bit for
    bit instead of the library's own (more accurate) float64 composition: the two differ by ~1.5e-5 px at the DTU
    shape, 100x the fp32 rounding of everything else on the path."""
    ref_inv = torch.inverse(proj[:, 0])
    out = torch.empty_like(proj)
    out[:, 0] = torch.eye(4, dtype=proj.dtype, device=proj.device)
    for v in range(1, proj.shape[1]):
        out[:, v] = torch.matmul(proj[:, v], ref_inv)
    return out


def release_workspaces(stream=None):
    """Drop every cached scratch buffer (they are re-created on demand), or only those of one CUDA stream."""
    if stream is None:
        _WS_CACHE.clear()
        return
    sid = stream.cuda_stream
    for key in [k for k in _WS_CACHE if k[2] == sid]:
        del _WS_CACHE[key]


# ------------------------------------------------------------------------------------------------
# (a2) homo_warping                                             reference models/module.py:96-139
# ------------------------------------------------------------------------------------------------
def homo_warping_fwd(src_fea, src_proj, ref_proj, depth_values):
    src_fea = _prep(src_fea, "src_fea", 4)
    src_proj = _prep(src_proj, "src_proj", 3)
    ref_proj = _prep(ref_proj, "ref_proj", 3)
    depth_values = _prep(depth_values, "depth_values", 2)
    B, C, H, W = src_fea.shape
    D = depth_values.shape[1]
    if src_proj.shape != (B, 4, 4) or ref_proj.shape != (B, 4, 4) or depth_values.shape[0] != B:
        raise RuntimeError("homo_warping: batch/shape mismatch: src_fea %s src_proj %s ref_proj %s depth_values %s"
                           % (tuple(src_fea.shape), tuple(src_proj.shape), tuple(ref_proj.shape),
                              tuple(depth_values.shape)))
    out = torch.empty((B, C, D, H, W), dtype=torch.float32, device=src_fea.device)
    with torch.cuda.device(src_fea.device):
        rc = _lib.load().mvs_homo_warping(_ptr(src_fea), _ptr(src_proj), _ptr(ref_proj), _ptr(depth_values), _ptr(out),
                                          B, C, D, H, W, _stream(src_fea))
    _lib.check(rc, "mvs_homo_warping")
    return out


def homo_warping_bwd(grad_out, src_proj, ref_proj, depth_values, H, W):
    grad_out = _prep(grad_out, "grad_out", 5)
    B, C, D = grad_out.shape[:3]
    grad_src = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        rc = _lib.load().mvs_homo_warping_bwd(_ptr(grad_out), _ptr(src_proj), _ptr(ref_proj), _ptr(depth_values),
                                              _ptr(grad_src), B, C, D, H, W, _stream(grad_out))
    _lib.check(rc, "mvs_homo_warping_bwd")
    return grad_src


class _HomoWarping(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src_fea, src_proj, ref_proj, depth_values):
        sp, rp, dv = (_prep(src_proj, "src_proj", 3), _prep(ref_proj, "ref_proj", 3),
                      _prep(depth_values, "depth_values", 2))
        ctx.save_for_backward(sp, rp, dv)
        ctx.hw = src_fea.shape[2:]
        return homo_warping_fwd(src_fea, sp, rp, dv)

    @staticmethod
    def backward(ctx, grad_out):
        sp, rp, dv = ctx.saved_tensors
        # gradient flows to the features only: the grid is built under no_grad in the reference (module.py:106)
        return homo_warping_bwd(grad_out, sp, rp, dv, ctx.hw[0], ctx.hw[1]), None, None, None


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    if isinstance(src_proj, torch.Tensor) and src_proj.is_cuda and src_proj.dtype == torch.float32:
        # module.py:107 with the reference's own calls (see compose_like_reference)
        src_proj = torch.matmul(src_proj, torch.inverse(ref_proj))
        ref_proj = torch.eye(4, dtype=src_proj.dtype, device=src_proj.device).expand_as(src_proj).contiguous()
    if torch.is_grad_enabled() and src_fea.requires_grad:
        return _HomoWarping.apply(src_fea, src_proj, ref_proj, depth_values)
    return homo_warping_fwd(src_fea, src_proj, ref_proj, depth_values)


# ------------------------------------------------------------------------------------------------
# (a2+a3, a8) fused warp + variance                              reference models/mvsnet.py:145-177
# ------------------------------------------------------------------------------------------------
def warp_variance_fwd(fea, proj, depth_values):
    fea = _prep(fea, "features", 5)
    proj = _prep(proj, "proj_matrices", 4)
    depth_values = _prep(depth_values, "depth_values", 2)
    B, V, C, H, W = fea.shape
    D = depth_values.shape[1]
    if proj.shape != (B, V, 4, 4):
        raise RuntimeError("Different number of images and projection matrices: features %s proj %s"
                           % (tuple(fea.shape), tuple(proj.shape)))
    if depth_values.shape[0] != B:
        raise RuntimeError("depth_values batch %d != %d" % (depth_values.shape[0], B))
    lib = _lib.load()
    var = torch.empty((B, C, D, H, W), dtype=torch.float32, device=fea.device)
    ws = _ws(lib.mvs_warp_variance_workspace_bytes(B, V, C, H, W), fea.device, "warp")
    with torch.cuda.device(fea.device), _timed("warp_variance_fwd", fea.device):
        rc = lib.mvs_warp_variance_fwd(_ptr(fea), _ptr(proj), _ptr(depth_values), _ptr(var), _ptr(ws), B, V, C, D, H, W,
                                       _stream(fea))
    _lib.check(rc, "mvs_warp_variance_fwd")
    return var


def warp_variance_bwd(grad_var, fea, proj, depth_values):
    grad_var = _prep(grad_var, "grad_var", 5)
    B, V, C, H, W = fea.shape
    D = depth_values.shape[1]
    lib = _lib.load()
    grad_fea = torch.empty_like(fea)
    ws = _ws(lib.mvs_warp_variance_bwd_workspace_bytes(B, V, C, H, W), fea.device)
    with torch.cuda.device(fea.device), _timed("warp_variance_bwd", fea.device):
        rc = lib.mvs_warp_variance_bwd(_ptr(grad_var), _ptr(fea), _ptr(proj), _ptr(depth_values), _ptr(grad_fea),
                                       _ptr(ws), B, V, C, D, H, W, _stream(fea))
    _lib.check(rc, "mvs_warp_variance_bwd")
    return grad_fea


class _WarpVariance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fea, proj, depth_values):
        f, p, dv = _prep(fea, "features", 5), _prep(proj, "proj_matrices", 4), _prep(depth_values, "depth_values", 2)
        ctx.save_for_backward(f, p, dv)  # features only: no warped volume is kept alive for backward
        return warp_variance_fwd(f, p, dv)

    @staticmethod
    def backward(ctx, grad_var):
        f, p, dv = ctx.saved_tensors
        return warp_variance_bwd(grad_var, f, p, dv), None, None


def warp_variance(fea, proj, depth_values):
    """fea [B,V,32,h,w] (view 0 = reference view) -> variance cost volume [B,32,D,h,w]."""
    if isinstance(proj, torch.Tensor) and proj.is_cuda and proj.dtype == torch.float32 and proj.dim() == 4:
        proj = compose_like_reference(proj)
    if torch.is_grad_enabled() and fea.requires_grad:
        return _WarpVariance.apply(fea, proj, depth_values)
    return warp_variance_fwd(fea, proj, depth_values)


# ------------------------------------------------------------------------------------------------
# (a4) CostRegNet                                    reference models/mvsnet.py:33-73, module.py:26-33
# ------------------------------------------------------------------------------------------------
def conv3d_bn_relu(x, w_folded, shift, relu=True, stride=1, tensor_cores=False):
    x = _prep(x, "x", 5)
    w_folded = _prep(w_folded, "weight", 5)
    shift = _prep(shift, "shift", 1)
    B, Cin, D, H, W = x.shape
    Cout = w_folded.shape[0]
    if w_folded.shape != (Cout, Cin, 3, 3, 3) or shift.shape[0] != Cout:
        raise RuntimeError("conv3d: weight %s / shift %s do not match input %s" % (tuple(w_folded.shape),
                                                                                   tuple(shift.shape), tuple(x.shape)))
    Do, Ho, Wo = [(n - 1) // stride + 1 for n in (D, H, W)]
    y = torch.empty((B, Cout, Do, Ho, Wo), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        fn = _lib.load().mvs_conv3d_bn_relu_tc if tensor_cores else _lib.load().mvs_conv3d_bn_relu
        rc = fn(_ptr(x), _ptr(w_folded), _ptr(shift), int(relu), _ptr(y), B, Cin, Cout, D, H, W, stride, _stream(x))
    _lib.check(rc, "mvs_conv3d_bn_relu")
    return y


def conv_transpose3d_bn_relu(x, w_folded, shift, relu=True, skip=None, tensor_cores=False):
    x = _prep(x, "x", 5)
    w_folded = _prep(w_folded, "weight", 5)
    shift = _prep(shift, "shift", 1)
    B, Cin, D, H, W = x.shape
    Cout = w_folded.shape[1]
    if w_folded.shape != (Cin, Cout, 3, 3, 3) or shift.shape[0] != Cout:
        raise RuntimeError("conv_transpose3d: weight %s does not match input %s" % (tuple(w_folded.shape),
                                                                                    tuple(x.shape)))
    if skip is not None:
        skip = _prep(skip, "skip", 5)
        if skip.shape != (B, Cout, 2 * D, 2 * H, 2 * W):
            raise RuntimeError("conv_transpose3d: skip shape %s != output shape" % (tuple(skip.shape),))
    y = torch.empty((B, Cout, 2 * D, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        fn = _lib.load().mvs_conv_transpose3d_bn_relu_tc if tensor_cores else _lib.load().mvs_conv_transpose3d_bn_relu
        rc = fn(_ptr(x), _ptr(w_folded), _ptr(shift), int(relu), _ptr(skip), _ptr(y), B, Cin, Cout, D, H, W, _stream(x))
    _lib.check(rc, "mvs_conv_transpose3d_bn_relu")
    return y


def cost_regularization(volume, folded, precision="fp32"):
    """volume [B,32,D,h,w]; folded = list of 11 (weight, shift) CUDA tensors in the order
    conv0..conv6, conv7, conv9, conv11, prob  ->  logits [B,D,h,w]."""
    volume = _prep(volume, "volume", 5)
    B, C, D, H, W = volume.shape
    if C != 32:
        raise RuntimeError("cost_regularization expects 32 channels, got %d" % C)
    if len(folded) != _lib.COSTREG_LAYERS:
        raise RuntimeError("cost_regularization expects %d folded layers" % _lib.COSTREG_LAYERS)
    prec = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16}[precision]
    lib = _lib.load()
    params = _lib.CostRegParams()
    keep = []
    for i, (w, s) in enumerate(folded):
        w, s = _prep(w, "weight%d" % i), _prep(s, "shift%d" % i)
        keep += [w, s]
        params.w[i] = w.data_ptr()
        params.shift[i] = s.data_ptr()
    nbytes = lib.mvs_costreg_workspace_bytes(B, D, H, W, prec)
    if nbytes == 0:
        raise RuntimeError("CostRegNet needs D, H, W divisible by 8 (got D=%d H=%d W=%d)" % (D, H, W))
    ws = _ws(nbytes, volume.device, "costreg")
    logits = torch.empty((B, D, H, W), dtype=torch.float32, device=volume.device)
    with torch.cuda.device(volume.device):
        rc = lib.mvs_costreg_fwd(_ptr(volume), ctypes.byref(params), _ptr(logits), _ptr(ws), B, D, H, W, prec,
                                 _stream(volume))
    _lib.check(rc, "mvs_costreg_fwd")
    return logits


class PreparedParams:
    """A (weight, shift) list checked once and packed into the C parameter struct: the models cache one per folded
    weight set, so that a forward pass does not re-validate ~40 tensors (0.15 ms of host time per depth map)."""

    def __init__(self, folded, kind):
        self.kind = kind
        if kind == "costreg":
            self.params, self.keep = _costreg_params(list(folded))
        else:
            self.params, self.keep = _featurenet_params(list(folded))
        self.device = self.keep[0].device


def _featurenet_params(folded):
    if len(folded) != _lib.FEATURENET_LAYERS:
        raise RuntimeError("FeatureNet expects %d folded layers" % _lib.FEATURENET_LAYERS)
    params = _lib.FeatureNetParams()
    keep = []
    for i, (w, s) in enumerate(folded):
        w = _prep(w, "featurenet weight %d" % i)
        s = _prep(s, "featurenet shift %d" % i, 1)
        keep += [w, s]
        params.w[i] = w.data_ptr()
        params.shift[i] = s.data_ptr()
    return params, keep


def _costreg_params(folded):
    if isinstance(folded, PreparedParams):
        if folded.kind != "costreg":
            raise RuntimeError("prepared FeatureNet parameters passed where CostRegNet's are expected")
        return folded.params, folded.keep
    if len(folded) != _lib.COSTREG_LAYERS:
        raise RuntimeError("cost_regularization expects %d folded layers" % _lib.COSTREG_LAYERS)
    params = _lib.CostRegParams()
    keep = []
    for i, (w, s) in enumerate(folded):
        w, s = _prep(w, "weight%d" % i), _prep(s, "shift%d" % i)
        keep += [w, s]
        params.w[i] = w.data_ptr()
        params.shift[i] = s.data_ptr()
    return params, keep


class Rcp8Features:
    """fp16 features of all views in the row-chunk-planar layout [B*V][h][4][w][8] written by the tensor-core
    FeatureNet (ops.featurenet_tc) and sampled by the fused warp kernel's TMA windows."""

    def __init__(self, data, B, V, h, w):
        self.data, self.B, self.V, self.h, self.w = data, B, V, h, w
        self.device = data.device

    def to_nchw(self):
        """[B,V,32,h,w] fp32 (tests / diagnostics)."""
        t = self.data.view(self.B, self.V, self.h, 4, self.w, 8).permute(0, 1, 3, 5, 2, 4)
        return t.reshape(self.B, self.V, 32, self.h, self.w).float()


def conv2d_bn_relu_tc(x, w_folded, shift, relu=True, stride=1, s2d_out=False):
    """One ConvBnReLU (reference models/module.py:8-15) with folded BN on the tensor-core kernel, fp16 operands.
    3x3 stride 1 or 5x5 stride 2; fp32 NCHW in / out (s2d_out: the space-to-depth form [N,4*Cout,H/2,W/2])."""
    x = _prep(x, "x", 4)
    w_folded = _prep(w_folded, "weight", 4)
    shift = _prep(shift, "shift", 1)
    N, Cin, H, W = x.shape
    Cout, k = w_folded.shape[0], w_folded.shape[2]
    if w_folded.shape != (Cout, Cin, k, k) or shift.shape[0] != Cout:
        raise RuntimeError("conv2d: weight %s / shift %s do not match input %s" % (tuple(w_folded.shape),
                                                                                   tuple(shift.shape), tuple(x.shape)))
    Ho, Wo = (H // stride, W // stride)
    shape = (N, 4 * Cout, Ho // 2, Wo // 2) if s2d_out else (N, Cout, Ho, Wo)
    y = torch.empty(shape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().mvs_conv2d_bn_relu_tc(_ptr(x), _ptr(w_folded), _ptr(shift), int(relu), _ptr(y), N, Cin, Cout, H, W,
                                               k, stride, int(s2d_out), _stream(x))
    _lib.check(rc, "mvs_conv2d_bn_relu_tc")
    return y


def featurenet_tc(imgs, folded, out=None):
    """FeatureNet.forward (reference models/mvsnet.py:10-30, eval mode) on the tensor cores.
    imgs [B,V,3,H,W] fp32; folded = 8 (weight, shift) CUDA tensors in layer order (native shapes, BN folded)
    -> Rcp8Features (fp16 [B*V][H/4][4][W/4][8]).  out: optional contiguous fp16 tensor of that shape to write into
    (a slice of a feature pool, see warp_variance_costreg_pool)."""
    u8 = isinstance(imgs, torch.Tensor) and imgs.dtype == torch.uint8
    if u8:  # 8-bit images as decoded from disk: /255 happens on the device (see mvs_featurenet_tc_fwd_u8)
        if not imgs.is_cuda or imgs.dim() != 5:
            raise RuntimeError("uint8 imgs must be a CUDA tensor [B,V,3,H,W], got %s on %s" % (tuple(imgs.shape), imgs.device))
        imgs = imgs.detach().contiguous()
    else:
        imgs = _prep(imgs, "imgs", 5)
    B, V, C, H, W = imgs.shape
    if C != 3:
        raise RuntimeError("FeatureNet expects 3-channel images, got %d" % C)
    lib = _lib.load()
    nbytes = lib.mvs_featurenet_tc_workspace_bytes(B * V, H, W)
    if nbytes == 0:
        raise RuntimeError("FeatureNet needs H, W divisible by 4 (got %dx%d)" % (H, W))
    if isinstance(folded, PreparedParams):
        if folded.kind != "featurenet":
            raise RuntimeError("prepared CostRegNet parameters passed where FeatureNet's are expected")
        params, keep = folded.params, folded.keep
    else:
        params, keep = _featurenet_params(folded)
    if keep[0].device != imgs.device:
        raise RuntimeError("FeatureNet weights are on %s, images on %s" % (keep[0].device, imgs.device))
    ws = _ws(nbytes, imgs.device, "featurenet")
    shape = (B * V, H // 4, 4, W // 4, 8)
    if out is None:
        out = torch.empty(shape, dtype=torch.float16, device=imgs.device)
    elif (tuple(out.shape) != shape or out.dtype != torch.float16 or out.device != imgs.device or not out.is_contiguous()):
        raise RuntimeError("featurenet_tc: out must be a contiguous fp16 tensor %s on %s" % (shape, imgs.device))
    with torch.cuda.device(imgs.device):
        fn = lib.mvs_featurenet_tc_fwd_u8 if u8 else lib.mvs_featurenet_tc_fwd
        rc = fn(_ptr(imgs), ctypes.byref(params), _ptr(out), _ptr(ws), B * V, H, W, _stream(imgs))
    _lib.check(rc, "mvs_featurenet_tc_fwd")
    return Rcp8Features(out, B, V, H // 4, W // 4)


def featurenet_fp32_supported(H, W):
    """The strict-fp32 FeatureNet kernels stage their halo tiles by TMA: every layer's row pitch (W, W/2, W/4 floats) must
    be a multiple of 16 bytes."""
    return H % 4 == 0 and W % 16 == 0


def featurenet_fp32(imgs, folded):
    """FeatureNet.forward (reference models/mvsnet.py:10-30, eval mode) at the reference's precision: fp32 FMA on the CUDA
    cores.  imgs [B,V,3,H,W] fp32; folded = 8 (weight, shift) CUDA tensors in layer order (native shapes, BN folded) or the
    PreparedParams of them -> fp32 features [B,V,32,H/4,W/4]."""
    imgs = _prep(imgs, "imgs", 5)
    B, V, C, H, W = imgs.shape
    if C != 3:
        raise RuntimeError("FeatureNet expects 3-channel images, got %d" % C)
    lib = _lib.load()
    nbytes = lib.mvs_featurenet_workspace_bytes(B * V, H, W)
    if nbytes == 0:
        raise RuntimeError("strict-fp32 FeatureNet needs H %% 4 == 0 and W %% 16 == 0 (got %dx%d)" % (H, W))
    if isinstance(folded, PreparedParams):
        if folded.kind != "featurenet":
            raise RuntimeError("prepared CostRegNet parameters passed where FeatureNet's are expected")
        params, keep = folded.params, folded.keep
    else:
        params, keep = _featurenet_params(folded)
    if keep[0].device != imgs.device:
        raise RuntimeError("FeatureNet weights are on %s, images on %s" % (keep[0].device, imgs.device))
    ws = _ws(nbytes, imgs.device, "featurenet32")
    out = torch.empty((B, V, 32, H // 4, W // 4), dtype=torch.float32, device=imgs.device)
    with torch.cuda.device(imgs.device):
        rc = lib.mvs_featurenet_fwd(_ptr(imgs), ctypes.byref(params), _ptr(out), _ptr(ws), B * V, H, W, _stream(imgs))
    _lib.check(rc, "mvs_featurenet_fwd")
    return out


def conv2d_bn_relu(x, w_folded, shift, relu=True, stride=1):
    """One ConvBnReLU (reference models/module.py:8-15, BN folded) on the strict-fp32 kernel: k3 s1 p1 or k5 s2 p2."""
    x = _prep(x, "x", 4)
    w_folded = _prep(w_folded, "weight", 4)
    shift = _prep(shift, "shift", 1)
    N, Cin, H, W = x.shape
    Cout, k = w_folded.shape[0], w_folded.shape[2]
    if w_folded.shape != (Cout, Cin, k, k) or shift.shape[0] != Cout:
        raise RuntimeError("conv2d: weight %s / shift %s do not match input %s" % (tuple(w_folded.shape), tuple(shift.shape),
                                                                                   tuple(x.shape)))
    y = torch.empty((N, Cout, (H - 1) // stride + 1, (W - 1) // stride + 1), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().mvs_conv2d_bn_relu(_ptr(x), _ptr(w_folded), _ptr(shift), int(relu), _ptr(y), N, Cin, Cout, H, W, k,
                                            stride, _stream(x))
    _lib.check(rc, "mvs_conv2d_bn_relu")
    return y


def warp_variance_cp8(fea, proj, depth_values):
    """Fused warp+variance with the 16-bit chunk-planar output: returns an fp16 tensor [B, 4, D, h, w, 8]
    (channel = chunk*8 + last index).  fea: fp32 [B,V,32,h,w] (exact fp32 arithmetic) or fp16 channels-last
    [B,V,h,w,32] (fp16 texels).  Mostly for tests/diagnostics; the model uses warp_variance_costreg_bf16."""
    lib = _lib.load()
    rcp8 = isinstance(fea, Rcp8Features)
    half_nhwc = (not rcp8) and fea.dtype == torch.float16
    if rcp8:
        B, V, H, W, C = fea.B, fea.V, fea.h, fea.w, 32
        fn, fea = lib.mvs_warp_variance_fwd_cp8_feat, fea.data
    elif half_nhwc:
        fea = fea.detach().contiguous()
        B, V, H, W, C = fea.shape
        fn = lib.mvs_warp_variance_fwd_cp8_f16
    else:
        fea = _prep(fea, "features", 5)
        B, V, C, H, W = fea.shape
        fn = lib.mvs_warp_variance_fwd_cp8
    proj = _prep(proj, "proj_matrices", 4)
    depth_values = _prep(depth_values, "depth_values", 2)
    D = depth_values.shape[1]
    vol = torch.empty((B, 4, D, H, W, 8), dtype=torch.float16, device=fea.device)
    ws1 = _ws(lib.mvs_warp_variance_workspace_bytes(B, V, C, H, W), fea.device)
    with torch.cuda.device(fea.device):
        rc = fn(_ptr(fea), _ptr(proj), _ptr(depth_values), _ptr(vol), _ptr(ws1), B, V, C, D, H, W, _stream(fea))
    _lib.check(rc, "mvs_warp_variance_fwd_cp8")
    return vol


def warp_variance_costreg_bf16(fea, proj, depth_values, folded, marks=None):
    """Tensor-core precision mode: fused warp+variance writing the fp16 CP8 volume, then the tcgen05 CostRegNet.
    fea: fp32 [B,V,32,h,w], or fp16 channels-last [B,V,h,w,32] (then no layout pre-pass runs) -> logits
    [B,D,h,w] (fp32).  `marks`, if a callable, is invoked between the two kernel families (stage timing)."""
    lib = _lib.load()
    rcp8 = isinstance(fea, Rcp8Features)
    half_nhwc = isinstance(fea, torch.Tensor) and fea.dtype == torch.float16
    if rcp8:
        B, V, H, W, C = fea.B, fea.V, fea.h, fea.w, 32
        fea = fea.data
    elif half_nhwc:
        if not fea.is_cuda or fea.dim() != 5 or fea.shape[-1] != 32 or not fea.is_contiguous():
            raise RuntimeError("fp16 features must be a contiguous CUDA tensor [B,V,h,w,32], got %s" % (tuple(fea.shape),))
        fea = fea.detach()
        B, V, H, W, C = fea.shape
    else:
        fea = _prep(fea, "features", 5)
        B, V, C, H, W = fea.shape
    proj = _prep(proj, "proj_matrices", 4)
    depth_values = _prep(depth_values, "depth_values", 2)
    D = depth_values.shape[1]
    if proj.shape != (B, V, 4, 4):
        raise RuntimeError("Different number of images and projection matrices: features %s proj %s"
                           % (tuple(fea.shape), tuple(proj.shape)))
    params, keep = _costreg_params(folded)
    nbytes = lib.mvs_costreg_workspace_bytes(B, D, H, W, _lib.PRECISION_BF16)
    if nbytes == 0:
        raise RuntimeError("CostRegNet needs D, H, W divisible by 8 (got D=%d H=%d W=%d)" % (D, H, W))
    vol = _ws(lib.mvs_volume_cp8_bytes(B, D, H, W), fea.device, "vol_cp8")
    ws1 = _ws(lib.mvs_warp_variance_workspace_bytes(B, V, C, H, W), fea.device, "warp")
    ws2 = _ws(nbytes, fea.device, "costreg")
    logits = torch.empty((B, D, H, W), dtype=torch.float32, device=fea.device)
    with torch.cuda.device(fea.device):
        fn = lib.mvs_warp_variance_fwd_cp8_feat if rcp8 else (
            lib.mvs_warp_variance_fwd_cp8_f16 if half_nhwc else lib.mvs_warp_variance_fwd_cp8)
        rc = fn(_ptr(fea), _ptr(proj), _ptr(depth_values), _ptr(vol), _ptr(ws1), B, V, C, D, H, W, _stream(fea))
        _lib.check(rc, "mvs_warp_variance_fwd_cp8")
        if marks is not None:
            marks("warp_variance")
        rc = lib.mvs_costreg_fwd_cp8(_ptr(vol), ctypes.byref(params), _ptr(logits), _ptr(ws2), B, D, H, W, _stream(fea))
        _lib.check(rc, "mvs_costreg_fwd_cp8")
    return logits


def warp_variance_costreg_pool(pool, view_ids, proj, depth_values, folded, marks=None):
    """Scan-level form of warp_variance_costreg_bf16: `pool` is an fp16 tensor [n_pool, h, 4, w, 8] of FeatureNet
    outputs (ops.featurenet_tc(..., out=pool[i:j])), `view_ids` the pool index of every view of this depth map (view 0 =
    reference view), proj [1,V,4,4], depth_values [1,D] -> logits [1,D,h,w]."""
    lib = _lib.load()
    if not (isinstance(pool, torch.Tensor) and pool.is_cuda and pool.dtype == torch.float16 and pool.dim() == 5 and
            pool.shape[2] == 4 and pool.shape[4] == 8 and pool.is_contiguous()):
        raise RuntimeError("pool must be a contiguous CUDA fp16 tensor [n, h, 4, w, 8]")
    n_pool, H, _, W, _ = pool.shape
    view_ids = [int(i) for i in view_ids]
    V = len(view_ids)
    proj = _prep(proj, "proj_matrices", 4)
    depth_values = _prep(depth_values, "depth_values", 2)
    if proj.shape != (1, V, 4, 4):
        raise RuntimeError("Different number of images and projection matrices: %d view ids, proj %s" % (V, tuple(proj.shape)))
    if depth_values.shape[0] != 1:
        raise RuntimeError("pool mode runs one reference view per call (depth_values %s)" % (tuple(depth_values.shape),))
    D = depth_values.shape[1]
    params, keep = _costreg_params(folded)
    nbytes = lib.mvs_costreg_workspace_bytes(1, D, H, W, _lib.PRECISION_BF16)
    if nbytes == 0:
        raise RuntimeError("CostRegNet needs D, H, W divisible by 8 (got D=%d H=%d W=%d)" % (D, H, W))
    vol = _ws(lib.mvs_volume_cp8_bytes(1, D, H, W), pool.device, "vol_cp8")
    ws1 = _ws(lib.mvs_warp_variance_workspace_bytes(1, V, 32, H, W), pool.device, "warp")
    ws2 = _ws(nbytes, pool.device, "costreg")
    logits = torch.empty((1, D, H, W), dtype=torch.float32, device=pool.device)
    ids = (ctypes.c_int * V)(*view_ids)
    with torch.cuda.device(pool.device):
        rc = lib.mvs_warp_variance_fwd_cp8_pool(_ptr(pool), n_pool, ids, _ptr(proj), _ptr(depth_values), _ptr(vol), _ptr(ws1),
                                                1, V, 32, D, H, W, _stream(pool))
        _lib.check(rc, "mvs_warp_variance_fwd_cp8_pool")
        if marks is not None:
            marks("warp_variance")
        rc = lib.mvs_costreg_fwd_cp8(_ptr(vol), ctypes.byref(params), _ptr(logits), _ptr(ws2), 1, D, H, W, _stream(pool))
        _lib.check(rc, "mvs_costreg_fwd_cp8")
    return logits


# ------------------------------------------------------------------------------------------------
# (a5-a7) softmax + depth + confidence     reference models/mvsnet.py:192-218, module.py:144-147
# ------------------------------------------------------------------------------------------------
def softmax_depth_conf(logits, depth_values, want_prob=False):
    logits = _prep(logits, "logits", 4)
    depth_values = _prep(depth_values, "depth_values", 2)
    B, D, H, W = logits.shape
    if depth_values.shape != (B, D):
        raise RuntimeError("depth_values %s does not match logits %s" % (tuple(depth_values.shape), tuple(logits.shape)))
    depth = torch.empty((B, H, W), dtype=torch.float32, device=logits.device)
    conf = torch.empty((B, H, W), dtype=torch.float32, device=logits.device)
    prob = torch.empty_like(logits) if want_prob else None
    with torch.cuda.device(logits.device):
        rc = _lib.load().mvs_softmax_depth_conf(_ptr(logits), _ptr(depth_values), _ptr(depth), _ptr(conf), _ptr(prob), B,
                                                D, H, W, _stream(logits))
    _lib.check(rc, "mvs_softmax_depth_conf")
    return (depth, conf, prob) if want_prob else (depth, conf)


def depth_regression_fwd(p, depth_values):
    p = _prep(p, "p", 4)
    depth_values = _prep(depth_values, "depth_values")
    B, D, H, W = p.shape
    if depth_values.dim() == 1 and depth_values.shape[0] == D:
        stride = 0
    elif depth_values.dim() == 2 and depth_values.shape == (B, D):
        stride = D
    else:
        raise RuntimeError("depth_values must be [B,D] or [D]; got %s for p %s" % (tuple(depth_values.shape),
                                                                                tuple(p.shape)))
    out = torch.empty((B, H, W), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        rc = _lib.load().mvs_depth_regression(_ptr(p), _ptr(depth_values), stride, _ptr(out), B, D, H, W, _stream(p))
    _lib.check(rc, "mvs_depth_regression")
    return out


class _DepthRegression(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, depth_values):
        ctx.save_for_backward(depth_values)
        return depth_regression_fwd(p, depth_values)

    @staticmethod
    def backward(ctx, g):
        (dv,) = ctx.saved_tensors
        dvb = dv.view(1, -1, 1, 1) if dv.dim() == 1 else dv.view(dv.shape[0], dv.shape[1], 1, 1)
        return g.unsqueeze(1) * dvb, None


def depth_regression(p, depth_values):
    if torch.is_grad_enabled() and p.requires_grad:
        return _DepthRegression.apply(p, depth_values)
    return depth_regression_fwd(p, depth_values)
