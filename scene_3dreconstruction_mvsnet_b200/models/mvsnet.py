"""Mirror of the reference's models/mvsnet.py: MVSNet.forward(imgs, proj_matrices, depth_values).

Same constructor, same forward signature, same output dict and the same parameter / buffer names
as the reference (so `load_state_dict` of a reference checkpoint is strict-clean), but the body of
the forward is the B200 path:

    FeatureNet (tcgen05 / fp32 CUDA-core kernels; cuDNN in training)  reference mvsnet.py:125
 -> fused plane-sweep warp + variance volume   [mvs_warp_variance_fwd]      :145-177
 -> CostRegNet as fused conv+BN+ReLU(+skip) kernels [mvs_costreg_fwd]        :180
 -> softmax + depth expectation + confidence  [mvs_softmax_depth_conf]      :192-218

In train() mode (or whenever autograd is recording) BatchNorm needs batch statistics and autograd
needs a graph, so CostRegNet and the softmax run as nn.Modules / torch ops on cuDNN, and only the
fused warp+variance op (forward and backward kernels) is ours -- exactly the split north_star asks
for.  The per-view warped volumes are never materialised in either mode.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .module import ConvBnReLU, ConvBnReLU3D, depth_regression, fold_bn, homo_warping  # noqa: F401


def _weights_key(module):
    """Cheap change detector for the folded-weight caches: autograd's version counter of every parameter and buffer
    (in-place updates: optimizer steps, load_state_dict) and the storage address of the first one (replaced storage:
    .to() / .cuda() go through nn.Module._apply, which MVSNet overrides to drop the caches as well).  The tensor list
    is collected once: ~10 us per forward instead of ~100 us for (data_ptr, version) of ~100 tensors."""
    ts = module.__dict__.get("_key_tensors")
    # (nn.DataParallel replicas copy __dict__ shallowly: a list that belongs to another module's parameters is rebuilt)
    if ts is None or ts[0] is not next(module.parameters()):
        ts = list(module.parameters()) + list(module.buffers())
        module.__dict__["_key_tensors"] = ts
    return (ts[0].data_ptr(), ts[-1].data_ptr()) + tuple([t._version for t in ts])


class FeatureNet(nn.Module):
    """8-layer 2-D CNN, 3 -> 32 channels at 1/4 resolution (reference mvsnet.py:10-30)."""

    def __init__(self):
        super().__init__()
        self.inplanes = 32
        self.conv0 = ConvBnReLU(3, 8, 3, 1, 1)
        self.conv1 = ConvBnReLU(8, 8, 3, 1, 1)
        self.conv2 = ConvBnReLU(8, 16, 5, 2, 2)
        self.conv3 = ConvBnReLU(16, 16, 3, 1, 1)
        self.conv4 = ConvBnReLU(16, 16, 3, 1, 1)
        self.conv5 = ConvBnReLU(16, 32, 5, 2, 2)
        self.conv6 = ConvBnReLU(32, 32, 3, 1, 1)
        self.feature = nn.Conv2d(32, 32, 3, 1, 1)

    def forward(self, x):
        x = self.conv1(self.conv0(x))
        x = self.conv4(self.conv3(self.conv2(x)))
        return self.feature(self.conv6(self.conv5(x)))

    _ORDER = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6")

    def folded_params(self):
        """[(weight NHWC, bias, stride, padding)] x 7 with eval-mode BN folded in; cached like CostRegNet's."""
        key = _weights_key(self)
        if getattr(self, "_folded", None) is None or key != self._folded_key:
            with torch.no_grad():
                out = []
                for n in self._ORDER:
                    layer = getattr(self, n)
                    w, b = fold_bn(layer.conv.weight, layer.bn, out_dim=0)
                    out.append((w.contiguous(memory_format=torch.channels_last), b, layer.conv.stride,
                                layer.conv.padding))
            self._folded, self._folded_key = out, key
        return self._folded

    def folded_native(self):
        """[(weight [Cout,Cin,k,k] fp32, shift)] x 8 in layer order with eval-mode BN folded in (the last entry is the
        plain `feature` conv and its bias): the parameter set of ops.featurenet_tc.  Cached like folded_params()."""
        key = _weights_key(self)
        if getattr(self, "_native", None) is None or key != self._native_key:
            with torch.no_grad():
                out = []
                for n in self._ORDER:
                    layer = getattr(self, n)
                    w, b = fold_bn(layer.conv.weight, layer.bn, out_dim=0)
                    out.append((w.float().contiguous(), b.float().contiguous()))
                out.append((self.feature.weight.detach().float().contiguous(), self.feature.bias.detach().float().contiguous()))
            self._native, self._native_key = out, key
            self._native_prepared = None
            ops.weights_changed()
        return self._native

    def native_prepared(self):
        """folded_native() validated once and packed for the C ABI (ops.PreparedParams)."""
        folded = self.folded_native()
        if getattr(self, "_native_prepared", None) is None:
            self._native_prepared = ops.PreparedParams(folded, "featurenet")
        return self._native_prepared

    def infer_half(self, x):
        """fp16 variant of infer() for the tensor-core mode: x fp16 channels-last -> fp16 channels-last features
        (1.34 ms vs 1.64 ms with TF32 at 5 x 1152x1600, same 1e-4 accuracy class), emitted in exactly the texel layout
        the fused warp+variance kernel samples, so no conversion pass runs in between."""
        key = _weights_key(self)
        if getattr(self, "_half", None) is None or key != self._half_key:
            with torch.no_grad():
                layers = [(w.half().contiguous(memory_format=torch.channels_last), b.half(), s, p)
                          for w, b, s, p in self.folded_params()]
                last = (self.feature.weight.detach().half().contiguous(memory_format=torch.channels_last),
                        self.feature.bias.detach().half())
            self._half, self._half_key = (layers, last), key
        layers, last = self._half
        for w, b, stride, padding in layers:
            x = torch.cudnn_convolution_relu(x, w, b, stride, padding, (1, 1), 1)
        return F.conv2d(x, last[0], last[1], 1, 1)

    def infer(self, x):
        """Inference path (still cuDNN -- FeatureNet is outside this build's scope): BN folded into the weights and
        cuDNN's fused conv+bias+ReLU, i.e. one kernel per layer instead of conv, BN and ReLU passes
        (measured at 5 x 1152x1600 on B200: 3.0 -> 1.6 ms with TF32, 9.6 -> 8.7 ms in strict fp32)."""
        for w, b, stride, padding in self.folded_params():
            x = torch.cudnn_convolution_relu(x, w, b, stride, padding, (1, 1), 1)
        return F.conv2d(x, self.feature.weight, self.feature.bias, 1, 1)


def _up(cin, cout):
    return nn.Sequential(
        nn.ConvTranspose3d(cin, cout, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False),
        nn.BatchNorm3d(cout), nn.ReLU(inplace=True))


class CostRegNet(nn.Module):
    """3-D conv U-Net (reference mvsnet.py:33-73)."""

    _ORDER = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6")

    def __init__(self):
        super().__init__()
        self.conv0 = ConvBnReLU3D(32, 8)
        self.conv1 = ConvBnReLU3D(8, 16, stride=2)
        self.conv2 = ConvBnReLU3D(16, 16)
        self.conv3 = ConvBnReLU3D(16, 32, stride=2)
        self.conv4 = ConvBnReLU3D(32, 32)
        self.conv5 = ConvBnReLU3D(32, 64, stride=2)
        self.conv6 = ConvBnReLU3D(64, 64)
        self.conv7 = _up(64, 32)
        self.conv9 = _up(32, 16)
        self.conv11 = _up(16, 8)
        self.prob = nn.Conv3d(8, 1, 3, stride=1, padding=1)
        self._folded = None
        self._folded_key = None

    def forward(self, x):
        """Autograd / training path (cuDNN)."""
        conv0 = self.conv0(x)
        conv2 = self.conv2(self.conv1(conv0))
        conv4 = self.conv4(self.conv3(conv2))
        x = self.conv6(self.conv5(conv4))
        x = conv4 + self.conv7(x)
        x = conv2 + self.conv9(x)
        x = conv0 + self.conv11(x)
        return self.prob(x)

    def folded_params(self):
        """[(weight, shift)] x 11 with eval-mode BN folded in; cached until a parameter or buffer
        is modified in place or replaced (load_state_dict, optimizer step, .to())."""
        key = _weights_key(self)
        if self._folded is None or key != self._folded_key:
            with torch.no_grad():
                out = [getattr(self, n).folded() for n in self._ORDER]
                for n in ("conv7", "conv9", "conv11"):
                    seq = getattr(self, n)
                    out.append(fold_bn(seq[0].weight, seq[1], out_dim=1))
                out.append((self.prob.weight.detach().contiguous(), self.prob.bias.detach().contiguous()))
            self._folded, self._folded_key = out, key
            self._prepared = None
            ops.weights_changed()
        return self._folded

    def folded_prepared(self):
        """folded_params() validated once and packed for the C ABI (ops.PreparedParams)."""
        folded = self.folded_params()
        if getattr(self, "_prepared", None) is None:
            self._prepared = ops.PreparedParams(folded, "costreg")
        return self._prepared

    def infer(self, volume, precision="fp32"):
        """Inference path: 11 fused CUDA launches, logits [B,D,h,w]."""
        return ops.cost_regularization(volume, self.folded_params(), precision)


class MVSNet(nn.Module):
    """Drop-in for the reference MVSNet (mvsnet.py:91-239).

    refine: kept for constructor compatibility.  The reference's RefineNet is dead code (F.cat does
            not exist, mvsnet.py:85; eval hard-codes refine=False, eval.py:308); refine=True raises.
    debug:  the reference's cv2.imshow bitmask; accepted and ignored (needs a display).
    precision: "fp32" (default) - everything in fp32 FMA arithmetic, matches the reference to fp32 rounding;
               "bf16" - CostRegNet on the tcgen05 tensor cores (bf16 operands, fp32 accumulate), cost volume stored
                        as bf16; the fused warp+variance kernel samples fp16 texels with packed-half interpolation
                        and accumulates Sum / Sum^2 in fp32 (TMA-window kernel, csrc/warp_variance_win.cu);
                        FeatureNet on cuDNN with TF32 allowed (PyTorch's default, i.e. what the reference itself
                        does on a GPU);
               "fast" - "bf16" with packed-half sums of deviations from the reference view in the fused warp kernel
                        (7 % faster kernel; large variances -- mismatched voxels -- up to 2^-6 relative error instead
                        of 2^-7); with featurenet="cudnn" also a cuDNN FeatureNet in fp16.
    featurenet: "auto" (default) - in the tensor-core modes FeatureNet runs on the same tcgen05 implicit-GEMM
                        kernel as CostRegNet (fp16 operands, fp32 accumulate, ops.featurenet_tc) and writes the warp
                        kernel's texel layout directly, and in "fp32" mode on the strict fp32 CUDA-core kernels
                        (ops.featurenet_fp32; image widths that are not a multiple of 16 fall back to cuDNN);
                        "cudnn" keeps it on cuDNN (always the case in train() mode and under autograd).
    """

    def __init__(self, refine=True, debug=0, precision="fp32", featurenet="auto"):
        super().__init__()
        if precision not in ("fp32", "bf16", "fast"):
            raise ValueError("precision must be 'fp32', 'bf16' or 'fast', got %r" % (precision,))
        if featurenet not in ("auto", "tc", "cudnn"):
            raise ValueError("featurenet must be 'auto', 'tc' or 'cudnn', got %r" % (featurenet,))
        self.featurenet = featurenet
        self.refine = refine
        self.debug = debug
        self.precision = precision
        self.feature = FeatureNet()
        self.cost_regularization = CostRegNet()
        self.stage_events = None  # set to a list to collect per-stage CUDA events (bench.py roofline legs)
        if self.refine:
            raise NotImplementedError(
                "refine=True: the reference RefineNet cannot run (F.cat, mvsnet.py:85); use refine=False like "
                "eval.py:308 and scripts/train_DTU.sh")

    def invalidate_folded(self):
        """Drop every derived copy of the weights: the BN-folded tensors cached on the modules and the packed 16-bit
        operand blocks the native library caches per weight pointer.  The caches notice replaced tensors and in-place
        updates that go through autograd's version counter (load_state_dict, optimizer steps, .to()); they cannot
        notice writes through `.data` (p.data.copy_(...), EMA / SWA weight swaps, manual initialisation), which change
        neither the pointer nor the version.  Call this after such an update and before the next eval forward."""
        for m in (self.feature, self.cost_regularization):
            for attr in ("_folded", "_native", "_half", "_prepared", "_native_prepared"):
                if hasattr(m, attr):
                    setattr(m, attr, None)
            m.__dict__.pop("_key_tensors", None)
        ops.weights_changed()

    def _apply(self, fn, *args, **kwargs):
        """.to() / .cuda() / .half() replace the parameters' storage: drop everything derived from the old one."""
        before = [(p.data_ptr(), p.dtype, p.device) for p in (next(self.feature.parameters()), next(self.cost_regularization.parameters()))]
        out = super()._apply(fn, *args, **kwargs)
        after = [(p.data_ptr(), p.dtype, p.device) for p in (next(self.feature.parameters()), next(self.cost_regularization.parameters()))]
        if before != after:   # a no-op .to(device) keeps the caches (and any CUDA graph captured over them)
            self.invalidate_folded()
        return out

    # -- feature extraction ----------------------------------------------------------------------
    def extract_features(self, imgs):
        """imgs [B,V,3,H,W] -> [B,V,32,H/4,W/4].  Eval: one batched pass over all views (identical
        math with running BN statistics).  Train: one pass per view, like the reference
        (mvsnet.py:125), because train-mode BN statistics are per call."""
        B, V = imgs.shape[:2]
        if self.training:
            return torch.stack([self.feature(img) for img in torch.unbind(imgs, 1)], 1)
        # channels_last suits cuDNN better for these tiny channel counts (measured on B200 at 5 x 1152x1600:
        # NCHW fp32 11.7 ms, NHWC fp32 9.8 ms, NHWC with TF32 allowed 3.0 ms)
        x = imgs.reshape(B * V, *imgs.shape[2:]).contiguous(memory_format=torch.channels_last)
        if self.precision in ("bf16", "fast"):
            # tensor-core modes: let cuDNN use its TF32 tensor-core kernels for FeatureNet (PyTorch's own
            # default for convolutions, i.e. what the reference does on a GPU); fp32 mode keeps the ambient setting
            with torch.backends.cudnn.flags(enabled=True, benchmark=torch.backends.cudnn.benchmark, allow_tf32=True):
                f = self._features_eval(x)
        else:
            f = self._features_eval(x)
        return f.view(B, V, *f.shape[1:])

    def extract_features_fp32(self, imgs):
        """Strict-fp32 inference: imgs [B,V,3,H,W] -> fp32 features [B,V,32,H/4,W/4] at the reference's precision on our own
        kernels (fp32 FMA, BN folded: ops.featurenet_fp32).  Shapes whose rows the TMA tensor maps cannot describe (W not a
        multiple of 16) and featurenet="cudnn" take the cuDNN path."""
        if self.featurenet != "cudnn" and ops.featurenet_fp32_supported(imgs.shape[-2], imgs.shape[-1]):
            return ops.featurenet_fp32(imgs.float(), self.feature.native_prepared())
        return self.extract_features(imgs)

    def extract_features_half(self, imgs):
        """Tensor-core mode: imgs [B,V,3,H,W] -> fp16 channels-last features [B,V,H/4,W/4,32]."""
        B, V = imgs.shape[:2]
        x = imgs.reshape(B * V, *imgs.shape[2:]).to(dtype=torch.float16, memory_format=torch.channels_last)
        f = self.feature.infer_half(x)                       # [B*V,32,h,w], channels_last strides
        f = f.permute(0, 2, 3, 1)                            # [B*V,h,w,32] view
        if not f.is_contiguous():
            f = f.contiguous()
        return f.view(B, V, *f.shape[1:])

    def _features_eval(self, x):
        if torch.is_grad_enabled():
            return self.feature(x).contiguous()
        return self.feature.infer(x).contiguous()

    def forward(self, imgs, proj_matrices, depth_values):
        if imgs.shape[1] != proj_matrices.shape[1]:
            raise AssertionError("Different number of images and projection matrices")  # mvsnet.py:106
        if not imgs.is_cuda:
            raise RuntimeError("MVSNet (B200 build) runs on CUDA devices only; there is no CPU fallback")
        marks = [] if self.stage_events is not None else None

        def mark(name):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(torch.cuda.current_stream(imgs.device))
                marks.append((name, e))

        mark("start")
        infer = (not torch.is_grad_enabled()) and (not self.training)
        tc_features = infer and self.precision in ("bf16", "fast") and self.featurenet != "cudnn"
        if imgs.dtype == torch.uint8 and not tc_features:
            # 8-bit images: the reference loader's normalisation (datasets/data_io.py).  A true fp32 division like
            # numpy's -- dividing by a Python scalar on CUDA multiplies by the rounded reciprocal instead.
            imgs = imgs.float() / torch.full((), 255.0, dtype=torch.float32, device=imgs.device)
        if tc_features:
            fea = ops.featurenet_tc(imgs if imgs.dtype == torch.uint8 else imgs.float(), self.feature.native_prepared())
        elif infer and self.precision == "fp32":
            fea = self.extract_features_fp32(imgs)
        elif infer and self.precision == "fast":
            fea = self.extract_features_half(imgs)
        else:
            fea = self.extract_features(imgs)
        mark("features")
        proj_matrices = proj_matrices.float()
        depth_values = depth_values.float()
        if not torch.is_grad_enabled() and not self.training and self.precision in ("bf16", "fast"):
            # tensor-core modes: the cost volume goes from the fused warp+variance kernel to the tcgen05 CostRegNet
            # as bf16 chunk-planar data; no fp32 volume is written
            logits = ops.warp_variance_costreg_bf16(fea, proj_matrices.float(), depth_values.float(),
                                                    self.cost_regularization.folded_prepared(), marks=mark)
            mark("cost_regularization")
            depth, photometric_confidence = ops.softmax_depth_conf(logits, depth_values)
            mark("depth_tail")
            if marks is not None:
                self.stage_events.append(marks)
            return {"depth": depth, "photometric_confidence": photometric_confidence}
        volume_variance = ops.warp_variance(fea, proj_matrices, depth_values)
        mark("warp_variance")

        if not torch.is_grad_enabled():
            if self.training:  # train-mode BN under no_grad (e.g. BN re-calibration): keep nn.Module semantics
                logits = self.cost_regularization(volume_variance).squeeze(1)
            else:
                logits = self.cost_regularization.infer(volume_variance, "fp32" if self.precision == "fp32" else "bf16")
            mark("cost_regularization")
            depth, photometric_confidence = ops.softmax_depth_conf(logits, depth_values)
            mark("depth_tail")
            if marks is not None:
                self.stage_events.append(marks)
            return {"depth": depth, "photometric_confidence": photometric_confidence}

        # autograd path (training): graph through cuDNN CostRegNet and torch softmax, ours elsewhere
        logits = self.cost_regularization(volume_variance).squeeze(1)
        prob_volume = F.softmax(logits, dim=1)
        depth = depth_regression(prob_volume, depth_values)
        with torch.no_grad():
            _, photometric_confidence = ops.softmax_depth_conf(logits, depth_values)
        return {"depth": depth, "photometric_confidence": photometric_confidence}


    # -- scan-level inference: features of an image are extracted once and shared by every reference view -----------
    def features_to_pool(self, imgs, out):
        """imgs [n,3,H,W] (fp32 in [0,1] or uint8) -> FeatureNet features written into `out`, a contiguous slice
        [n, H/4, 4, W/4, 8] of an fp16 feature pool (tensor-core modes, eval, no autograd)."""
        if self.training or torch.is_grad_enabled() or self.precision not in ("bf16", "fast"):
            raise RuntimeError("features_to_pool: tensor-core inference only (eval mode, no_grad, precision 'bf16')")
        x = imgs if imgs.dtype == torch.uint8 else imgs.float()
        ops.featurenet_tc(x.unsqueeze(0), self.feature.native_prepared(), out=out)

    def forward_from_pool(self, pool, view_ids, proj_matrices, depth_values):
        """The rest of forward() for one reference view whose views' features are pool[view_ids] (view_ids[0] = the
        reference view; mvsnet.py:126-236 without :125).  proj_matrices [1,V,4,4], depth_values [1,D]."""
        if len(view_ids) != proj_matrices.shape[1]:
            raise AssertionError("Different number of images and projection matrices")  # mvsnet.py:106
        if self.training or torch.is_grad_enabled() or self.precision not in ("bf16", "fast"):
            raise RuntimeError("forward_from_pool: tensor-core inference only (eval mode, no_grad, precision 'bf16')")
        marks = [] if self.stage_events is not None else None

        def mark(name):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(torch.cuda.current_stream(pool.device))
                marks.append((name, e))

        mark("start")
        depth_values = depth_values.float()
        logits = ops.warp_variance_costreg_pool(pool, view_ids, proj_matrices.float(), depth_values,
                                                self.cost_regularization.folded_prepared(), marks=mark)
        mark("cost_regularization")
        depth, photometric_confidence = ops.softmax_depth_conf(logits, depth_values)
        mark("depth_tail")
        if marks is not None:
            self.stage_events.append(marks)
        return {"depth": depth, "photometric_confidence": photometric_confidence}


def mvsnet_loss(depth_est, depth_gt, mask):
    """Smooth-L1 over valid pixels (reference mvsnet.py:242-244)."""
    mask = mask > 0.5
    return F.smooth_l1_loss(depth_est[mask], depth_gt[mask], reduction="mean")
