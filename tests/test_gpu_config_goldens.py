"""GPU parity at the BASELINE.json shapes against outputs of the UNMODIFIED reference (tests/golden/config_*.npz,
made by `python tests/make_golden.py --config-sized` in the build container, calibrated checkpoint).

  fp32 mode        north_star's bar: depth max abs <= 1e-3 x depth range, confidence <= 1e-4 relative.
  tensor-core mode ONE stated tolerance for all three shapes (TC_DEPTH_TOL, TC_CONF_TOL below), for the depth map AND
                   the photometric confidence.  The confidence is checked twice: (a) on EVERY pixel as the 4-plane
                   probability sum gathered at the REFERENCE's index, so that a flip of trunc(index) cannot
                   hide a probability error, and (b) the model's own output wherever its index equals the reference's;
                   the fraction of pixels whose index flipped is bounded too (TC_FLIP_FRAC).
The stage samples (features, variance volume and logits at sampled pixels) pin the fused kernels to the reference at
full size as well.
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden
from scene_3dreconstruction_mvsnet_b200 import ops, synth
from scene_3dreconstruction_mvsnet_b200.models import MVSNet

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

FP32_DEPTH_TOL = 1e-3      # x depth range, max abs      (north_star)
FP32_CONF_REL = 1e-4       # relative                    (north_star)
TC_DEPTH_TOL = 5e-3        # x depth range, max abs      (tensor-core mode, stated in DESIGN.md section 2)
TC_DEPTH_MEAN_TOL = 5e-4   # x depth range, mean abs
TC_CONF_TOL = 2e-2         # absolute, confidence in [0, 1]
TC_CONF_MEAN_TOL = 1e-3
TC_LOGITS_TOL = 8e-2       # absolute, max over 256 sampled pixels x D planes (logits std ~0.8); the reference itself on this
TC_LOGITS_RMS_TOL = 1.5e-2  # GPU with PyTorch defaults (TF32 convolutions) is off by the same amount (profiles/r02_precision.md)
TC_FLIP_FRAC = 0.05        # pixels whose trunc(index expectation) differs from the reference's
SAFE_BAND = 2e-3           # fp32 mode: |index_f - round(index_f)| below which trunc(index) may flip (as in test_gpu_parity)
TAGS = ["c1", "c3", "c2"]


def _load(tag):
    g = load_golden("config_%s.npz" % tag)
    name = str(g["name"])
    imgs, proj, dv = synth.make_named(name, B=1, seed=0)
    assert hashlib.sha1(imgs.numpy().tobytes()).hexdigest() == str(g["imgs_sha1"]), "synthetic inputs differ from the golden run"
    return g, imgs.to(DEV), proj.to(DEV), dv.to(DEV)


def _model(weights, precision):
    m = MVSNet(refine=False, precision=precision)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    return m.to(DEV).eval()


def _sum4_at(prob, idx):
    """4-plane sum of mvsnet.py:216-218 gathered at a given integer index map.  prob [1,D,h,w], idx [1,h,w] int64."""
    p = torch.nn.functional.pad(prob, (0, 0, 0, 0, 1, 2))          # planes -1 .. D+1
    c = torch.cumsum(p, 1)
    c = torch.nn.functional.pad(c, (0, 0, 0, 0, 1, 0))             # c[k] = sum of padded planes < k
    hi = torch.gather(c, 1, (idx + 4).unsqueeze(1))
    lo = torch.gather(c, 1, idx.unsqueeze(1))
    return (hi - lo).squeeze(1)


def measure(tag, weights, precision):
    """Error figures of one precision mode against the reference golden (also used by tools/precision_report.py).
    precision "reference_cuda_default": not ours -- the reference's own ATen calls (oracle/torch_port.py) on this GPU with
    PyTorch's default settings (cuDNN convolutions may use TF32), the yardstick for the tensor-core mode's tolerance."""
    g, imgs, proj, dv = _load(tag)
    rng = float(dv.max() - dv.min())
    if precision == "reference_cuda_default":
        from oracle import torch_port
        sd = {k: torch.from_numpy(v).to(DEV) for k, v in weights.items()}
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        try:
            st = {}
            out = torch_port.mvsnet_forward(imgs, proj, dv, sd, stages=st)
        finally:
            torch.backends.cudnn.allow_tf32 = old
    else:
        m = _model(weights, precision)
        with torch.no_grad():
            out = m(imgs, proj, dv)
    depth, conf = out["depth"].cpu().numpy(), out["photometric_confidence"].cpu().numpy()
    safe = np.abs(g["index_f"] - np.round(g["index_f"])) > SAFE_BAND
    r = {"depth_max": float(np.abs(depth - g["depth"]).max() / rng), "depth_mean": float(np.abs(depth - g["depth"]).mean() / rng),
         "conf_safe_max": float(np.abs(conf - g["conf"])[safe].max()), "conf_mean": float(np.abs(conf - g["conf"]).mean()),
         "conf_safe_rel": float((np.abs(conf - g["conf"]) / g["conf"])[safe].max()), "safe_frac": float(safe.mean())}
    if precision == "reference_cuda_default":
        yx = g["sample_yx"]
        lg = st["logits"][0][:, yx[:, 0], yx[:, 1]].cpu().numpy()
        r["logits_max"] = float(np.abs(lg - g["logits_samples"]).max())
        r["logits_rms"] = float(np.sqrt(((lg - g["logits_samples"]) ** 2).mean()))
        idx_ref = torch.from_numpy(np.trunc(g["index_f"]).astype(np.int64)).to(DEV)
        r["conf_at_ref_index_max"] = float(np.abs(_sum4_at(st["prob"], idx_ref).cpu().numpy() - g["conf"]).max())
    elif precision != "fp32":
        with torch.no_grad():
            fea = ops.featurenet_tc(imgs, m.feature.folded_native())
            logits = ops.warp_variance_costreg_bf16(fea, proj, dv, m.cost_regularization.folded_params())
            _, _, prob = ops.softmax_depth_conf(logits, dv, want_prob=True)
            vol = ops.warp_variance_cp8(fea, proj, dv)
        idx_ref = torch.from_numpy(np.trunc(g["index_f"]).astype(np.int64)).to(DEV)
        s4 = _sum4_at(prob, idx_ref).cpu().numpy()
        r["conf_at_ref_index_max"] = float(np.abs(s4 - g["conf"]).max())
        D = prob.shape[1]
        idx_own = (prob * torch.arange(D, dtype=torch.float32, device=DEV).view(1, D, 1, 1)).sum(1).long()
        same = (idx_own == idx_ref).cpu().numpy()
        r["index_flip_frac"] = float(1.0 - same.mean())
        r["conf_same_index_max"] = float(np.abs(conf - g["conf"])[same].max())
        yx = g["sample_yx"]
        f = fea.to_nchw()[0][:, :, yx[:, 0], yx[:, 1]].cpu().numpy()
        r["features_max"] = float(np.abs(f - g["features_samples"]).max() / np.abs(g["features_samples"]).max())
        lg = logits[0][:, yx[:, 0], yx[:, 1]].cpu().numpy()
        r["logits_max"] = float(np.abs(lg - g["logits_samples"]).max())
        r["logits_rms"] = float(np.sqrt(((lg - g["logits_samples"]) ** 2).mean()))
        v = vol[0].float().permute(0, 4, 1, 2, 3).reshape(32, *vol.shape[2:5])[:, :, yx[:16, 0], yx[:16, 1]].cpu().numpy()
        ref = g["variance_samples"]
        r["variance_max_scaled"] = float((np.abs(v - ref) / (np.abs(ref) * 2.0 ** -6 + 2e-2)).max())
    return r


@pytest.mark.parametrize("tag", TAGS)
def test_fp32_mode_matches_reference_at_config_size(tag, weights):
    r = measure(tag, weights, "fp32")
    assert r["depth_max"] < FP32_DEPTH_TOL, r
    assert r["conf_safe_rel"] < FP32_CONF_REL, r


@pytest.mark.parametrize("tag", TAGS)
def test_tensor_core_mode_matches_reference_at_config_size(tag, weights):
    r = measure(tag, weights, "bf16")
    assert r["depth_max"] < TC_DEPTH_TOL and r["depth_mean"] < TC_DEPTH_MEAN_TOL, r
    assert r["conf_at_ref_index_max"] < TC_CONF_TOL, r
    assert r["conf_same_index_max"] < TC_CONF_TOL and r["conf_mean"] < TC_CONF_MEAN_TOL, r
    assert r["index_flip_frac"] < TC_FLIP_FRAC, r
    assert r["features_max"] < 4e-3, r          # fp16 FeatureNet, fraction of max |feature|
    assert r["logits_max"] < TC_LOGITS_TOL and r["logits_rms"] < TC_LOGITS_RMS_TOL, r
    assert r["variance_max_scaled"] < 1.0, r     # |var - ref| <= 2^-6 |ref| + 2e-2 (includes the fp16 FeatureNet's error)
