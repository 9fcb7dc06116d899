"""Functional torch-CPU port of the reference forward pass (TEST INFRASTRUCTURE ONLY).

This is the second form of the oracle: where oracle/mvsnet_oracle.c restates the arithmetic
independently in C, this file issues the SAME PyTorch library calls the reference makes
(F.grid_sample, F.conv3d, F.conv_transpose3d, F.batch_norm, F.softmax, F.avg_pool3d,
torch.gather), written as stateless functions over a state_dict instead of nn.Modules.
It is what bench.py times as the CPU baseline (`--impl reference`, `cpu_baseline.kind="port"`):
the reference's CPU path *is* these ATen/oneDNN calls on the host cores, and /root/reference
itself does not exist on the GPU box.  bench.py also runs it with CUDA tensors as the
`cuda_eager_baseline` (the reference's stock eager-GPU path: the same calls dispatch to cuDNN / ATen CUDA).

Pinned against the unmodified reference in tests/test_oracle_golden.py (golden vectors made by
tests/make_golden.py, which imports /root/reference/models in the build container).

Citations are to /root/reference.
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


_TRAIN = [False]   # set by mvsnet_forward_train: BatchNorm with batch statistics (nn.Module.train())


def _bn(x, sd, p):
    if _TRAIN[0]:
        return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                            True, 0.1, EPS)
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        False, 0.0, EPS)


def feature_net(img, sd, pre="feature."):
    """FeatureNet.forward, eval mode (models/mvsnet.py:10-30)."""
    spec = [("conv0", 1, 1), ("conv1", 1, 1), ("conv2", 2, 2), ("conv3", 1, 1), ("conv4", 1, 1), ("conv5", 2, 2),
            ("conv6", 1, 1)]
    x = img
    for name, stride, pad in spec:
        x = F.conv2d(x, sd[f"{pre}{name}.conv.weight"], None, stride, pad)
        x = F.relu(_bn(x, sd, f"{pre}{name}.bn."))
    return F.conv2d(x, sd[pre + "feature.weight"], sd[pre + "feature.bias"], 1, 1)


def plane_sweep_grid(src_proj, ref_proj, depth_values, h, w):
    """Sampling grid of homo_warping (models/module.py:106-133): [B, D*h, w, 2], normalised for
    align_corners=True exactly as the reference does."""
    B, D = depth_values.shape
    P = src_proj @ torch.inverse(ref_proj)
    R, t = P[:, :3, :3], P[:, :3, 3:4]
    dev = src_proj.device  # CPU for the oracle; bench.py also runs this port on the GPU as the eager-CUDA baseline
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=dev),
                            torch.arange(w, dtype=torch.float32, device=dev), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, device=dev)))  # [3, hw]
    q = (R @ pix.unsqueeze(0).expand(B, -1, -1)).unsqueeze(2) * depth_values.view(B, 1, D, 1) + t.view(B, 3, 1, 1)
    xy = q[:, :2] / q[:, 2:3]
    gx = xy[:, 0] / ((w - 1) / 2) - 1
    gy = xy[:, 1] / ((h - 1) / 2) - 1
    return torch.stack((gx, gy), dim=3).view(B, D * h, w, 2)


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    """models/module.py:96-139 -- grid_sample is called WITHOUT align_corners (module.py:135)."""
    B, C, h, w = src_fea.shape
    D = depth_values.shape[1]
    grid = plane_sweep_grid(src_proj, ref_proj, depth_values, h, w)
    out = F.grid_sample(src_fea, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    return out.view(B, C, D, h, w)


def variance_volume(features, projs, depth_values):
    """models/mvsnet.py:145-177 (eval-mode in-place branch). features: list of [B,C,h,w]."""
    D = depth_values.shape[1]
    V = len(features)
    s = features[0].unsqueeze(2).repeat(1, 1, D, 1, 1)
    q = s ** 2
    for fea, proj in zip(features[1:], projs[1:]):
        wv = homo_warping(fea, proj, projs[0], depth_values)
        s += wv
        q += wv.pow_(2)
    return q.div_(V).sub_(s.div_(V).pow_(2))


def cost_reg_net(x, sd, pre="cost_regularization."):
    """CostRegNet.forward, eval mode (models/mvsnet.py:33-73)."""
    def cbr(x, name, stride=1):
        return F.relu(_bn(F.conv3d(x, sd[f"{pre}{name}.conv.weight"], None, stride, 1), sd, f"{pre}{name}.bn."))

    def up(x, name):
        y = F.conv_transpose3d(x, sd[f"{pre}{name}.0.weight"], None, stride=2, padding=1, output_padding=1)
        return F.relu(_bn(y, sd, f"{pre}{name}.1."))

    c0 = cbr(x, "conv0")
    c2 = cbr(cbr(c0, "conv1", 2), "conv2")
    c4 = cbr(cbr(c2, "conv3", 2), "conv4")
    x = cbr(cbr(c4, "conv5", 2), "conv6")
    x = c4 + up(x, "conv7")
    x = c2 + up(x, "conv9")
    x = c0 + up(x, "conv11")
    return F.conv3d(x, sd[pre + "prob.weight"], sd[pre + "prob.bias"], 1, 1)


def depth_tail(logits, depth_values):
    """models/mvsnet.py:192-218: softmax, depth expectation, 4-plane photometric confidence."""
    B, D = depth_values.shape
    p = F.softmax(logits, dim=1)
    depth = torch.sum(p * depth_values.view(B, D, 1, 1), 1)
    sum4 = 4 * F.avg_pool3d(F.pad(p.unsqueeze(1), (0, 0, 0, 0, 1, 2)), (4, 1, 1), stride=1, padding=0).squeeze(1)
    idx = torch.sum(p * torch.arange(D, dtype=torch.float32, device=p.device).view(1, D, 1, 1), 1).long()
    conf = torch.gather(sum4, 1, idx.unsqueeze(1)).squeeze(1)
    return depth, conf, p


@torch.no_grad()
def mvsnet_forward(imgs, proj_matrices, depth_values, sd, stages=None):
    """MVSNet.forward, eval mode, refine=False (models/mvsnet.py:103-236).
    imgs [B,V,3,H,W], proj_matrices [B,V,4,4], depth_values [B,D]; sd = reference state_dict.
    `stages`, if a dict, receives the intermediate tensors for stage-wise comparison."""
    views = torch.unbind(imgs, 1)
    projs = torch.unbind(proj_matrices, 1)
    assert len(views) == len(projs)
    feats = [feature_net(v, sd) for v in views]
    if stages is not None:
        stages["features"] = torch.stack(feats, 1).clone()
    var = variance_volume(feats, projs, depth_values)
    if stages is not None:
        stages["variance"] = var.clone()
    logits = cost_reg_net(var, sd).squeeze(1)
    depth, conf, prob = depth_tail(logits, depth_values)
    if stages is not None:
        stages["logits"] = logits
        stages["prob"] = prob
    return {"depth": depth, "photometric_confidence": conf}


def variance_volume_train(features, projs, depth_values):
    """models/mvsnet.py:145-169,177 (training branch: out-of-place sums, every warped volume stays alive for autograd)."""
    D = depth_values.shape[1]
    V = len(features)
    s = features[0].unsqueeze(2).repeat(1, 1, D, 1, 1)
    q = s ** 2
    for fea, proj in zip(features[1:], projs[1:]):
        wv = homo_warping(fea, proj, projs[0], depth_values)
        s = s + wv
        q = q + wv ** 2
    return q.div_(V).sub_(s.div_(V).pow_(2))


def mvsnet_forward_train(imgs, proj_matrices, depth_values, sd):
    """MVSNet.forward in train() mode with autograd (models/mvsnet.py:103-236; the step of train.py:241-300 around it is
    the caller's): BatchNorm on batch statistics, the out-of-place variance branch, depth by soft argmin."""
    _TRAIN[0] = True
    try:
        views = torch.unbind(imgs, 1)
        projs = torch.unbind(proj_matrices, 1)
        feats = [feature_net(v, sd) for v in views]
        var = variance_volume_train(feats, projs, depth_values)
        logits = cost_reg_net(var, sd).squeeze(1)
        p = F.softmax(logits, dim=1)
        B, D = depth_values.shape
        depth = torch.sum(p * depth_values.view(B, D, 1, 1), 1)
        with torch.no_grad():
            _, conf, _ = depth_tail(logits.detach(), depth_values)
    finally:
        _TRAIN[0] = False
    return {"depth": depth, "photometric_confidence": conf}
