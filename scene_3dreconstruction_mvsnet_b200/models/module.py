"""Mirror of the reference's models/module.py public surface, backed by libmvsnet_b200.so.

`homo_warping` and `depth_regression` keep the reference signatures (module.py:96, module.py:144)
and run hand-written sm_100a kernels through the C ABI; the layer containers keep the reference's
attribute names so that checkpoints load unchanged (state_dict keys `*.conv.weight`, `*.bn.*`).
"""
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


class ConvBnReLU(nn.Module):
    """2-D conv + BN + ReLU of FeatureNet (reference module.py:6-13).  FeatureNet is outside this
    build's scope and runs on cuDNN."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)), inplace=True)


class ConvBnReLU3D(nn.Module):
    """3-D conv + BN + ReLU of CostRegNet (reference module.py:26-33).  In eval/no-grad mode
    CostRegNet does not call this forward: it folds BN into the weights and runs the fused CUDA
    kernels (ops.cost_regularization).  The nn.Module path is the autograd/training path."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm3d(out_channels)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)), inplace=True)

    def folded(self):
        return fold_bn(self.conv.weight, self.bn, out_dim=0)


def fold_bn(weight, bn, out_dim):
    """Eval-mode BatchNorm folded into the preceding (bias-free) convolution:
    w' = w * gamma/sqrt(var+eps) along `out_dim`, shift = beta - mean*gamma/sqrt(var+eps)."""
    scale = bn.weight.detach() / (bn.running_var + bn.eps).sqrt()
    shape = [1] * weight.dim()
    shape[out_dim] = -1
    w = (weight.detach() * scale.view(shape)).contiguous()
    shift = (bn.bias.detach() - bn.running_mean * scale).contiguous()
    return w, shift


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    """Plane-sweep warp of a source feature map (reference module.py:96-139).
    src_fea [B,C,H,W], src_proj/ref_proj [B,4,4], depth_values [B,D] -> [B,C,D,H,W].
    Differentiable w.r.t. src_fea only, like the reference (grid built under no_grad)."""
    return ops.homo_warping(src_fea, src_proj, ref_proj, depth_values)


def depth_regression(p, depth_values):
    """sum_d p[:,d] * depth_values[:,d] (reference module.py:144-147); depth_values [B,D] or [D]."""
    return ops.depth_regression(p, depth_values)
