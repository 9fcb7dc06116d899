"""ctypes/numpy front end of oracle/mvsnet_oracle.c (TEST INFRASTRUCTURE ONLY).

Each wrapper mirrors one C function; see the C file for the reference file:line citations.
All arrays are C-contiguous float32 numpy arrays in the reference's layouts (NCHW / NCDHW).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmvsnet_oracle.so")
_lib = None

BN_EPS = 1e-5  # nn.BatchNorm3d default, models/module.py:30


def build(force=False):
    """Compile the C restatement with gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "mvsnet_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _opt(a):
    if a is None:
        return None, None
    return _f(a)


def num_threads():
    return int(lib().orc_num_threads())


def compose_homography(src_proj, ref_proj):
    """[B,4,4] x2 -> rot [B,3,3], trans [B,3]   (module.py:107-109)"""
    src_proj, _ = _f(src_proj)
    ref_proj, _ = _f(ref_proj)
    B = src_proj.shape[0]
    rot = np.empty((B, 3, 3), np.float32)
    trans = np.empty((B, 3), np.float32)
    for b in range(B):
        rc = lib().orc_compose_homography(
            src_proj[b].ctypes.data_as(ctypes.c_void_p), ref_proj[b].ctypes.data_as(ctypes.c_void_p),
            rot[b].ctypes.data_as(ctypes.c_void_p), trans[b].ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise ValueError("singular reference projection matrix")
    return rot, trans


def compose_all(proj_matrices):
    """proj [B,V,4,4] (view 0 = reference) -> rot [B,V-1,3,3], trans [B,V-1,3]"""
    proj_matrices = np.asarray(proj_matrices, np.float32)
    B, V = proj_matrices.shape[:2]
    rot = np.empty((B, V - 1, 3, 3), np.float32)
    trans = np.empty((B, V - 1, 3), np.float32)
    for v in range(1, V):
        r, t = compose_homography(proj_matrices[:, v], proj_matrices[:, 0])
        rot[:, v - 1], trans[:, v - 1] = r, t
    return rot, trans


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    """module.py:96-139"""
    src_fea, pf = _f(src_fea)
    depth_values, pd = _f(depth_values)
    B, C, H, W = src_fea.shape
    D = depth_values.shape[1]
    rot, trans = compose_homography(src_proj, ref_proj)
    out = np.empty((B, C, D, H, W), np.float32)
    lib().orc_homo_warp(pf, rot.ctypes.data_as(ctypes.c_void_p), trans.ctypes.data_as(ctypes.c_void_p), pd,
                        out.ctypes.data_as(ctypes.c_void_p), B, C, D, H, W)
    return out


def warp_variance(fea, proj_matrices, depth_values):
    """fea [B,V,C,H,W] (view 0 = ref), proj [B,V,4,4], depth [B,D] -> variance volume [B,C,D,H,W]
    (mvsnet.py:145-177)"""
    fea, pf = _f(fea)
    depth_values, pd = _f(depth_values)
    B, V, C, H, W = fea.shape
    D = depth_values.shape[1]
    rot, trans = compose_all(proj_matrices)
    out = np.empty((B, C, D, H, W), np.float32)
    lib().orc_warp_variance(pf, rot.ctypes.data_as(ctypes.c_void_p), trans.ctypes.data_as(ctypes.c_void_p), pd,
                            out.ctypes.data_as(ctypes.c_void_p), B, V, C, D, H, W)
    return out


def warp_variance_bwd(grad_var, fea, proj_matrices, depth_values):
    """-> grad_fea [B,V,C,H,W]"""
    grad_var, pg = _f(grad_var)
    fea, pf = _f(fea)
    depth_values, pd = _f(depth_values)
    B, V, C, H, W = fea.shape
    D = depth_values.shape[1]
    rot, trans = compose_all(proj_matrices)
    out = np.empty_like(fea)
    lib().orc_warp_variance_bwd(pg, pf, rot.ctypes.data_as(ctypes.c_void_p), trans.ctypes.data_as(ctypes.c_void_p),
                                pd, out.ctypes.data_as(ctypes.c_void_p), B, V, C, D, H, W)
    return out


def conv3d(x, weight, bias=None, bn=None, relu=False, stride=1):
    """bn = (gamma, beta, running_mean, running_var) or None   (module.py:26-33, mvsnet.py:62)"""
    x, px = _f(x)
    weight, pw = _f(weight)
    bias, pb = _opt(bias)
    B, Cin, D, H, W = x.shape
    Cout = weight.shape[0]
    assert weight.shape == (Cout, Cin, 3, 3, 3)
    bnp = [None] * 4
    keep = []
    if bn is not None:
        for i, t in enumerate(bn):
            a, p = _f(t)
            keep.append(a)
            bnp[i] = p
    Do, Ho, Wo = [(n - 1) // stride + 1 for n in (D, H, W)]
    out = np.empty((B, Cout, Do, Ho, Wo), np.float32)
    lib().orc_conv3d(px, pw, pb, bnp[0], bnp[1], bnp[2], bnp[3], ctypes.c_float(BN_EPS), int(relu),
                     out.ctypes.data_as(ctypes.c_void_p), B, Cin, Cout, D, H, W, stride)
    return out


def conv_transpose3d(x, weight, bn=None, relu=False, skip=None):
    """ConvTranspose3d(k3,s2,p1,op1,bias=False) [+BN eval] [+ReLU] [+skip]   (mvsnet.py:46-59,69-71)"""
    x, px = _f(x)
    weight, pw = _f(weight)
    skip, ps = _opt(skip)
    B, Cin, D, H, W = x.shape
    Cout = weight.shape[1]
    assert weight.shape == (Cin, Cout, 3, 3, 3)
    bnp = [None] * 4
    keep = []
    if bn is not None:
        for i, t in enumerate(bn):
            a, p = _f(t)
            keep.append(a)
            bnp[i] = p
    out = np.empty((B, Cout, 2 * D, 2 * H, 2 * W), np.float32)
    lib().orc_convT3d(px, pw, bnp[0], bnp[1], bnp[2], bnp[3], ctypes.c_float(BN_EPS), int(relu), ps,
                      out.ctypes.data_as(ctypes.c_void_p), B, Cin, Cout, D, H, W)
    return out


def cost_regularization(volume, sd, prefix="cost_regularization."):
    """CostRegNet.forward in eval mode (mvsnet.py:64-73) from a reference state_dict (numpy values)."""
    def g(k):
        return np.asarray(sd[prefix + k], np.float32)

    def cbr(x, name, stride=1):
        bn = tuple(g(f"{name}.bn.{s}") for s in ("weight", "bias", "running_mean", "running_var"))
        return conv3d(x, g(f"{name}.conv.weight"), None, bn, True, stride)

    def up(x, name, skip):
        bn = tuple(g(f"{name}.1.{s}") for s in ("weight", "bias", "running_mean", "running_var"))
        return conv_transpose3d(x, g(f"{name}.0.weight"), bn, True, skip)

    c0 = cbr(volume, "conv0")
    c2 = cbr(cbr(c0, "conv1", 2), "conv2")
    c4 = cbr(cbr(c2, "conv3", 2), "conv4")
    x = cbr(cbr(c4, "conv5", 2), "conv6")
    x = up(x, "conv7", c4)
    x = up(x, "conv9", c2)
    x = up(x, "conv11", c0)
    return conv3d(x, g("prob.weight"), g("prob.bias"), None, False, 1)


def softmax_depth_conf(logits, depth_values, want_prob=False):
    """logits [B,D,H,W] -> depth, conf, idx (float expectation of the index) [, prob]
    (mvsnet.py:192-193,204,214-218)"""
    logits, pl = _f(logits)
    depth_values, pd = _f(depth_values)
    B, D, H, W = logits.shape
    depth = np.empty((B, H, W), np.float32)
    conf = np.empty((B, H, W), np.float32)
    idx = np.empty((B, H, W), np.float32)
    prob = np.empty((B, D, H, W), np.float32) if want_prob else None
    lib().orc_softmax_depth_conf(pl, pd, prob.ctypes.data_as(ctypes.c_void_p) if want_prob else None,
                                 depth.ctypes.data_as(ctypes.c_void_p), conf.ctypes.data_as(ctypes.c_void_p),
                                 idx.ctypes.data_as(ctypes.c_void_p), B, D, H, W)
    return (depth, conf, idx, prob) if want_prob else (depth, conf, idx)


def depth_regression(p, depth_values):
    """module.py:144-147; depth_values [B,D] or [D]"""
    p, pp = _f(p)
    depth_values, pd = _f(depth_values)
    B, D, H, W = p.shape
    out = np.empty((B, H, W), np.float32)
    lib().orc_depth_regression(pp, pd, D if depth_values.ndim == 2 else 0, out.ctypes.data_as(ctypes.c_void_p),
                               B, D, H, W)
    return out
