"""GPU parity of the tcgen05 CostRegNet / FeatureNet path against the oracle.

Tolerance (stated separately from the fp32 path, as north_star allows): operands are rounded to fp16
(11-bit significand; bf16 until round 2), accumulation is fp32.  Per layer the oracle is evaluated on the SAME
fp16-rounded inputs and weights, so the remaining difference is accumulation order plus the fp16 rounding of the
stored output: |y - ref| <= 2^-10 |ref| + 2.5e-4.  End to end (11 layers) logits are compared at LOGITS_TOL absolute
on logits of std ~0.5, and the depth map at 5e-3 x depth range (tests/test_gpu_config_goldens.py holds the same bar
at the BASELINE shapes, confidence included)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from scene_3dreconstruction_mvsnet_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


LOGITS_TOL = 1e-2


def bf16_round(t):
    """Rounds to the 16-bit storage format of the tensor-core path (fp16 since round 2; the name is historical)."""
    return t.to(torch.float16).to(torch.float32)


def check(y, ref, what):
    y = y.cpu().numpy().astype(np.float64)
    err = np.abs(y - ref)
    tol = np.abs(ref) * 2.0 ** -10 + 2.5e-4
    bad = err > tol
    assert not bad.any(), "%s: %d/%d outside tolerance, max err %.4g at %s (ref %.4g got %.4g)" % (
        what, bad.sum(), bad.size, err.max(), np.unravel_index(err.argmax(), err.shape), ref.flat[err.argmax()],
        y.flat[err.argmax()])


CONV_CASES = [
    # cin, cout, stride, (D,H,W)
    (32, 8, 1, (6, 7, 40)),      # conv0-like, one tile, partial rows
    (32, 8, 1, (10, 24, 230)),   # several x/y tiles, z segments
    (16, 16, 1, (5, 9, 33)),
    (8, 1, 1, (9, 10, 50)),      # prob layer: paired taps, fp32 single-channel output
    (32, 32, 1, (4, 20, 27)),
    (64, 64, 1, (3, 9, 13)),     # two output-channel groups
    (8, 16, 2, (8, 12, 40)),     # stride 2 with paired taps across parity sub-planes
    (16, 32, 2, (6, 10, 44)),
    (32, 64, 2, (4, 8, 12)),
    # several items per CTA with odd segment lengths: the three epilogue warp sets of the kw-folded fold kernels take the
    # output planes in turn across item boundaries (accumulator-ring phases tracked for every block by every set)
    (32, 8, 1, (11, 96, 330)),
    (8, 1, 1, (11, 96, 330)),
    (16, 16, 1, (7, 64, 200)),
    (8, 1, 1, (5, 12, 20)),      # prob layer narrower than its fixed 30-column tile
    (32, 8, 1, (5, 9, 22)),      # conv0 likewise
    (32, 32, 1, (9, 40, 150)),   # conv4-like: depth fold with 32-column blocks, several tiles and segments
]


@pytest.mark.parametrize("cin,cout,stride,dims", CONV_CASES)
def test_tc_conv3d(cin, cout, stride, dims):
    g = torch.Generator().manual_seed(cin * 7 + cout + dims[2])
    x = bf16_round(torch.randn(1, cin, *dims, generator=g))
    w = bf16_round(torch.randn(cout, cin, 3, 3, 3, generator=g) * (1.0 / (27 * cin) ** 0.5))
    shift = torch.randn(cout, generator=g) * 0.3
    relu = cout != 1
    y = ops.conv3d_bn_relu(x.to(DEV), w.to(DEV), shift.to(DEV), relu=relu, stride=stride, tensor_cores=True)
    ref = orc.conv3d(x.numpy(), w.numpy(), shift.numpy(), None, relu, stride).astype(np.float64)
    assert tuple(y.shape) == ref.shape
    check(y, ref, "conv %d->%d s%d %s" % (cin, cout, stride, dims))


@pytest.mark.parametrize("cin,cout,dims,with_skip", [(16, 8, (4, 6, 30), True), (32, 16, (3, 5, 21), True),
                                                     (64, 32, (2, 4, 9), True), (16, 8, (5, 13, 120), False)])
def test_tc_conv_transpose3d(cin, cout, dims, with_skip):
    g = torch.Generator().manual_seed(cin + dims[2])
    x = bf16_round(torch.randn(1, cin, *dims, generator=g))
    w = bf16_round(torch.randn(cin, cout, 3, 3, 3, generator=g) * (1.0 / (8 * cin) ** 0.5))
    shift = torch.randn(cout, generator=g) * 0.3
    skip = bf16_round(torch.randn(1, cout, *[2 * d for d in dims], generator=g)) if with_skip else None
    y = ops.conv_transpose3d_bn_relu(x.to(DEV), w.to(DEV), shift.to(DEV), relu=True,
                                     skip=skip.to(DEV) if with_skip else None, tensor_cores=True)
    bn = (np.ones(cout, np.float32), shift.numpy(), np.zeros(cout, np.float32), np.full(cout, 1 - orc.BN_EPS, np.float32))
    ref = orc.conv_transpose3d(x.numpy(), w.numpy(), bn, True, skip.numpy() if with_skip else None).astype(np.float64)
    check(y, ref, "convT %d->%d %s" % (cin, cout, dims))


def _cp8_to_ncdhw(cp8):
    B, _, D, h, w, _ = cp8.shape
    return cp8.permute(0, 1, 5, 2, 3, 4).reshape(B, 32, D, h, w).float().cpu().numpy().astype(np.float64)


def _check_cp8_against_oracle(fea, proj, dv, what):
    """Both CP8 entries (fp32 NCHW features / fp16 channels-last features) sample fp16 texels of ALL views with
    packed-half interpolation, accumulate the deviations from the reference view and their squares in packed half and
    store fp16.  Oracle: the C restatement on fp16-rounded features.  Tolerance (stated for this kernel): 2^-6 |ref| + 8e-3
    (times (V-1)/7 beyond 7 source views) on N(0,1) features -- independent random features are the worst case of the deviation sums (every source view differs
    from the reference view by O(1), so squares of ~10-40 are summed with fp16 rounding), mean < 1.5e-3.  What the depth map
    needs is pinned separately against the reference's outputs (tests/test_gpu_config_goldens.py)."""
    fea_q = fea.half().float()
    ref = orc.warp_variance(fea_q.numpy(), proj.numpy(), dv.numpy()).astype(np.float64)
    fea16 = fea.to(DEV).half().permute(0, 1, 3, 4, 2).contiguous()             # [B,V,h,w,32] fp16 channels-last
    for name, arg in (("fp16 nhwc", fea16), ("fp32 nchw", fea.to(DEV))):
        back = _cp8_to_ncdhw(ops.warp_variance_cp8(arg, proj.to(DEV), dv.to(DEV)))
        err = np.abs(back - ref)
        # the packed-half sums round once per view: beyond 7 source views the bound grows with the view count
        grow = max(1.0, (fea.shape[1] - 1) / 7.0)
        bad = err > grow * (np.abs(ref) * 2.0 ** -6 + 8e-3)
        assert not bad.any(), "%s / %s: %d/%d outside tolerance, max err %.4g at %s" % (
            what, name, bad.sum(), bad.size, err.max(), np.unravel_index(err.argmax(), err.shape))
        assert err.mean() < 1.5e-3, "%s / %s: mean err %.4g" % (what, name, err.mean())


@pytest.mark.parametrize("B,V,h,w,D", [(1, 5, 24, 40, 16), (2, 3, 9, 33, 5), (1, 4, 16, 104, 8), (1, 2, 8, 31, 3),
                                       (1, 1, 8, 16, 4), (1, 3, 40, 72, 40),
                                       (1, 8, 16, 40, 8),    # 7 source views: run-time view loop, smaller windows
                                       (1, 10, 16, 40, 8)])  # 9 source views: more than windows -> every plane gathers
def test_warp_variance_cp8(B, V, h, w, D):
    """Tensor-core-mode fused kernel (TMA-window generation) on camera-like geometry."""
    from scene_3dreconstruction_mvsnet_b200 import synth
    fea = synth.make_features(B, V, 32, h, w, seed=V)
    _, proj, dv = synth.make_inputs(B=B, V=V, H=4 * h, W=4 * w, D=D, focal=0.9 * w, interval_scale=8.0, yaw=0.04, seed=D)
    _check_cp8_against_oracle(fea, proj, dv, "B%d V%d %dx%d D%d" % (B, V, h, w, D))


def _custom_proj(V, h, w, focal, poses):
    """poses: per source view (roll, yaw, zoom, tx, ty, tz); view 0 = identity camera."""
    import math
    K = np.array([[focal, 0, w / 2.0], [0, focal, h / 2.0], [0, 0, 1]], np.float64)
    out = np.zeros((1, V, 4, 4), np.float32)
    for v in range(V):
        roll, yaw, zoom, tx, ty, tz = poses[v - 1] if v else (0, 0, 1, 0, 0, 0)
        Rz = np.array([[math.cos(roll), -math.sin(roll), 0], [math.sin(roll), math.cos(roll), 0], [0, 0, 1]])
        Ry = np.array([[math.cos(yaw), 0, math.sin(yaw)], [0, 1, 0], [-math.sin(yaw), 0, math.cos(yaw)]])
        E = np.eye(4)
        E[:3, :3] = Rz @ Ry
        E[:3, 3] = (tx, ty, tz)
        Kv = K.copy()
        Kv[0, 0] *= zoom
        Kv[1, 1] *= zoom
        P = E.copy()
        P[:3, :4] = Kv @ E[:3, :4]
        out[0, v] = P
    return torch.from_numpy(out)


STRESS = {
    # in-plane rotation: epipolar lines and tile footprints are oblique (windows need extra rows, segments split)
    "roll30": dict(poses=[(0.5, 0.0, 1.0, 60, 10, 0), (-0.3, 0.05, 1.0, -80, 0, 0)], dv=(425, 20.0)),
    # 3x zoom-in source view: footprint of a 32-pixel row spans ~96 texels > window -> per-plane global gather
    "zoom3": dict(poses=[(0.0, 0.0, 3.0, 40, 0, 0), (0.1, 0.0, 0.4, -40, 5, 0)], dv=(425, 20.0)),
    # large depth steps: the footprint jumps by many texels per plane
    "bigstep": dict(poses=[(0.0, 0.02, 1.0, 300, 0, 0), (0.0, 0.0, 1.0, 0, -250, 0)], dv=(200, 150.0)),
    # source camera moved forward past the first planes: q_z changes sign inside the sweep (mirrored coordinates)
    "behind": dict(poses=[(0.0, 0.0, 1.0, 30, 0, -500.0), (0.2, 0.3, 1.0, 100, 0, -430.0)], dv=(400, 12.0)),
    # non-monotonic, partly non-positive depth hypotheses (the reference does not validate them)
    "weird_depths": dict(poses=[(0.0, 0.0, 1.0, 60, 0, 0), (0.0, 0.1, 1.0, -60, 0, 0)], dv=None),
}


@pytest.mark.parametrize("name", sorted(STRESS))
def test_warp_variance_cp8_stress_geometry(name):
    """Geometry the window planner must survive: every case is checked against the oracle, whichever mix of windows,
    split segments and global gathers the kernel chooses."""
    from scene_3dreconstruction_mvsnet_b200 import synth
    cfg = STRESS[name]
    V, h, w, D = 3, 24, 72, 12
    fea = synth.make_features(1, V, 32, h, w, seed=11)
    proj = _custom_proj(V, h, w, 0.9 * w, cfg["poses"])
    if cfg["dv"] is None:
        dv = torch.tensor([[500.0, 430.0, 900.0, 0.0, -300.0, 650.0, 425.0, 1e4, 0.5, 700.0, 640.0, 520.0]])
    else:
        dv = (cfg["dv"][0] + cfg["dv"][1] * torch.arange(D, dtype=torch.float32)).unsqueeze(0)
    _check_cp8_against_oracle(fea, proj, dv, name)


@pytest.mark.parametrize("precision", ["bf16", "fast"])
@pytest.mark.parametrize("case", ["case_a", "case_b"])
def test_tc_costreg_and_depth_golden(case, precision, request, weights):
    from test_gpu_parity import cu, load_model, maxabs
    c = request.getfixturevalue(case)
    m = load_model(weights, precision=precision)
    logits = m.cost_regularization.infer(cu(c["variance"]), "bf16")
    assert maxabs(logits, c["logits"]) < LOGITS_TOL
    with torch.no_grad():
        out = m(cu(c["imgs"]), cu(c["proj"]), cu(c["dv"]))
    rng = float(c["dv"].max() - c["dv"].min())
    assert maxabs(out["depth"], c["depth"]) < 5e-3 * rng


# ------------------------------------------------------------------------------------------------
# FeatureNet on the tensor-core kernel (fp16 operands, fp32 accumulate).  Floating-point kernel: the reference is
# torch's fp32 conv2d on the same fp16-rounded inputs and weights; tolerance = fp16 rounding of the stored output
# (2^-11 relative) plus accumulation order: |y - ref| <= 2^-10 |ref| + 2e-3.
# ------------------------------------------------------------------------------------------------
def f16_round(t):
    return t.half().float()


def check16(y, ref, what):
    err = (y.cpu().double() - ref.double()).abs()
    tol = ref.double().abs() * 2.0 ** -10 + 2e-3
    assert bool((err <= tol).all()), "%s: max err %.4g (ref absmax %.4g)" % (what, err.max().item(), ref.abs().max().item())


@pytest.mark.parametrize("cin,cout,k,hw,N", [(3, 8, 3, (20, 44), 2), (8, 8, 3, (37, 130), 1), (16, 16, 3, (24, 40), 3),
                                             (32, 32, 3, (9, 13), 2), (8, 16, 5, (24, 52), 2), (16, 32, 5, (20, 36), 1),
                                             (8, 16, 5, (6, 260), 1),
                                             # kw-folded variants (EPI=3: row pitch 32, N = 3*Cout), ragged widths / heights
                                             (16, 16, 3, (50, 95), 1), (16, 32, 5, (40, 136), 2), (8, 16, 5, (34, 180), 1),
                                             (16, 8, 3, (19, 61), 2)])
def test_tc_conv2d(cin, cout, k, hw, N):
    g = torch.Generator().manual_seed(cin * 31 + cout + hw[1])
    x = f16_round(torch.randn(N, cin, *hw, generator=g))
    w = f16_round(torch.randn(cout, cin, k, k, generator=g) * (1.0 / (k * k * cin) ** 0.5))
    shift = torch.randn(cout, generator=g) * 0.3
    stride = 2 if k == 5 else 1
    y = ops.conv2d_bn_relu_tc(x.to(DEV), w.to(DEV), shift.to(DEV), relu=True, stride=stride)
    ref = torch.relu(torch.nn.functional.conv2d(x, w, shift, stride=stride, padding=k // 2))
    assert tuple(y.shape) == tuple(ref.shape)
    check16(y, ref, "conv2d %d->%d k%d %s" % (cin, cout, k, hw))


def test_tc_conv2d_space_to_depth_output():
    """A layer that feeds a stride-2 layer writes [N, 4*C, H/2, W/2] with channel = ((y&1)*2 + (x&1))*C + c."""
    g = torch.Generator().manual_seed(5)
    x = f16_round(torch.randn(2, 8, 12, 40, generator=g))
    w = f16_round(torch.randn(8, 8, 3, 3, generator=g) * 0.1)
    shift = torch.randn(8, generator=g) * 0.3
    y = ops.conv2d_bn_relu_tc(x.to(DEV), w.to(DEV), shift.to(DEV), relu=True, s2d_out=True)
    ref = torch.relu(torch.nn.functional.conv2d(x, w, shift, padding=1))
    ref = torch.stack([ref[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], 1).reshape(2, 32, 6, 20)
    check16(y, ref, "s2d output")


@pytest.mark.parametrize("B,V,H,W", [(1, 3, 64, 96), (2, 2, 32, 160)])
def test_featurenet_tc_matches_torch(B, V, H, W, weights):
    """Whole FeatureNet (8 layers, fp16 activations) against the nn.Module in fp32 on the BN-calibrated checkpoint:
    features are O(1); 8 layers of fp16 storage give ~1e-3 absolute."""
    from test_gpu_parity import load_model
    m = load_model(weights, precision="bf16")
    g = torch.Generator().manual_seed(3)
    imgs = torch.rand(B, V, 3, H, W, generator=g).to(DEV)
    with torch.no_grad():
        fea = ops.featurenet_tc(imgs, m.feature.folded_native())
        ref = m.extract_features(imgs)                        # cuDNN, TF32 allowed in this mode
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            ref32 = torch.stack([m.feature(img) for img in torch.unbind(imgs, 1)], 1)
    got = fea.to_nchw()
    scale = ref32.abs().max().item()
    assert (got - ref32).abs().max().item() < 4e-3 * max(scale, 1.0), "max err %.4g, feature absmax %.4g" % (
        (got - ref32).abs().max().item(), scale)
    # and it is at least as close to the fp32 result as the TF32 cuDNN path of the same mode is, within a factor
    assert (got - ref32).abs().mean().item() < 5 * (ref - ref32).abs().mean().item() + 1e-4


def test_uint8_images_equal_float_images(weights):
    """8-bit images (as decoded from disk) give bit-identical depth maps to the float32 images the reference's loader
    produces from them (value / 255 in fp32), on the tensor-core path and -- through the torch fallback -- in fp32 mode."""
    from test_gpu_parity import load_model
    from scene_3dreconstruction_mvsnet_b200 import synth
    imgs, proj, dv = synth.make_inputs(B=1, V=3, H=64, W=96, D=16, focal=90.0, interval_scale=8.0, seed=2)
    u8 = (imgs * 255.0).round().to(torch.uint8)
    f32 = u8.float() / 255.0
    for precision in ("bf16", "fp32"):
        m = load_model(weights, precision=precision)
        with torch.no_grad():
            a = m(u8.to(DEV), proj.to(DEV), dv.to(DEV))
            b = m(f32.to(DEV), proj.to(DEV), dv.to(DEV))
        assert torch.equal(a["depth"], b["depth"]), "%s: max diff %g" % (precision, (a["depth"] - b["depth"]).abs().max().item())
        assert torch.equal(a["photometric_confidence"], b["photometric_confidence"]), precision


# ------------------------------------------------------------------------------------------------ full size (C2)
def _to_rcp8(fea):
    """[B,V,32,h,w] -> ops.Rcp8Features (fp16 [B*V][h][4][w][8]), the layout the tensor-core FeatureNet writes."""
    B, V, C, h, w = fea.shape
    t = fea.half().view(B * V, 4, 8, h, w).permute(0, 3, 1, 4, 2).contiguous()
    return ops.Rcp8Features(t, B, V, h, w)


@pytest.mark.parametrize("workload", ["c2_dtu_5view_1152x1600", "c2_dtu_5view_1152x1600_rot"])
def test_full_size_c2_window_kernel_matches_strict_kernel(workload):
    """DTU eval shape (5 views, 288x400 feature maps, D=192), rectified and rotated cameras: the TMA-window kernel
    against the strict fp32 kernel (itself pinned to the oracle at small sizes) on the same fp16-rounded features.
    Covers every tile / depth-chunk / window-segment / window-shape combination of the benchmark geometries (708 M
    voxels x channels).  Tolerance: the one stated for the tensor-core mode's fused kernel, 2^-6 |ref| + 8e-3 on N(0,1)
    features (packed-half deviation sums: large variances carry the fp16 rounding of d^2 and of their sum)."""
    from scene_3dreconstruction_mvsnet_b200 import synth
    B, V, h, w, D = 1, 5, 288, 400, 192
    fea = synth.make_features(B, V, 32, h, w, seed=4).half().float()
    _, proj, dv = synth.make_named(workload)
    cp8 = ops.warp_variance_cp8(_to_rcp8(fea.to(DEV)), proj.to(DEV), dv.to(DEV))
    ref = ops.warp_variance(fea.to(DEV), proj.to(DEV), dv.to(DEV))
    got = cp8.permute(0, 1, 5, 2, 3, 4).reshape(B, 32, D, h, w).float()
    err = (got - ref).abs()
    tol = ref.abs() * 2.0 ** -6 + 8e-3
    assert bool((err <= tol).all()), "max err %.4g, max err/tol %.3f" % (err.max().item(), (err / tol).max().item())
    assert err.mean().item() < 1.5e-3


def test_warp_variance_cp8_large_features_stay_finite():
    """Features ~100x larger than a trained FeatureNet emits: squared deviations overflow fp16 inside the kernel; the
    volume must saturate at the largest finite fp16 value, never inf or NaN (the last FeatureNet conv has no BN / ReLU,
    so a checkpoint does not bound its output)."""
    from scene_3dreconstruction_mvsnet_b200 import synth
    fea = synth.make_features(1, 3, 32, 16, 40, seed=9) * 150.0
    _, proj, dv = synth.make_inputs(B=1, V=3, H=64, W=160, D=8, focal=36.0, interval_scale=8.0, yaw=0.04, seed=2)
    vol = ops.warp_variance_cp8(_to_rcp8(fea.to(DEV)), proj.to(DEV), dv.to(DEV)).float()
    assert bool(torch.isfinite(vol).all()) and float(vol.min()) >= 0.0
    assert float(vol.max()) == 65504.0


def test_full_size_c2_forward_properties(weights):
    """DTU eval shape through MVSNet.forward in the tensor-core mode: deterministic, depth inside the hypothesis
    range, confidence in [0, 1], and close to the strict-fp32 mode within the stated tolerance."""
    from test_gpu_parity import load_model
    from scene_3dreconstruction_mvsnet_b200 import synth
    imgs, proj, dv = (t.to(DEV) for t in synth.make_named("c2_dtu_5view_1152x1600"))
    m = load_model(weights, precision="bf16")
    with torch.no_grad():
        a = m(imgs, proj, dv)
        b = m(imgs, proj, dv)
        s = load_model(weights, precision="fp32")(imgs, proj, dv)
    assert torch.equal(a["depth"], b["depth"]) and torch.equal(a["photometric_confidence"], b["photometric_confidence"])
    assert a["depth"].shape == (1, 288, 400)
    assert float(a["depth"].min()) >= float(dv.min()) - 1e-2 and float(a["depth"].max()) <= float(dv.max()) + 1e-2
    c = a["photometric_confidence"]
    assert float(c.min()) >= 0 and float(c.max()) <= 1 + 1e-5
    rng = float(dv.max() - dv.min())
    assert float((a["depth"] - s["depth"]).abs().mean()) < 5e-4 * rng
    assert float((a["depth"] - s["depth"]).abs().max()) < 5e-3 * rng   # the one tolerance stated for this mode


def test_packed_weight_cache_follows_weight_changes(weights):
    """The tensor-core layers cache their packed 16-bit weights per weight pointer; the models invalidate the cache
    whenever BatchNorm is re-folded.  In-place updates (an optimizer step) and load_state_dict must be picked up."""
    from test_gpu_parity import load_model
    from scene_3dreconstruction_mvsnet_b200 import synth
    imgs, proj, dv = synth.make_inputs(B=1, V=3, H=64, W=96, D=16, focal=90.0, interval_scale=8.0, seed=5)
    imgs, proj, dv = imgs.to(DEV), proj.to(DEV), dv.to(DEV)
    m = load_model(weights, precision="bf16")
    with torch.no_grad():
        a = m(imgs, proj, dv)["depth"].clone()
        m.cost_regularization.conv0.conv.weight.mul_(1.5)          # in place: same pointer, new version
        m.feature.conv1.conv.weight.mul_(0.5)
        b = m(imgs, proj, dv)["depth"].clone()
    fresh = load_model(weights, precision="bf16")
    with torch.no_grad():
        fresh.cost_regularization.conv0.conv.weight.mul_(1.5)
        fresh.feature.conv1.conv.weight.mul_(0.5)
        c = fresh(imgs, proj, dv)["depth"]
        m.load_state_dict(load_model(weights, precision="bf16").state_dict())
        d = m(imgs, proj, dv)["depth"]
    assert not torch.equal(a, b), "changed weights must change the result"
    assert torch.equal(b, c), "in-place update must give the same result as a fresh model with those weights"
    assert torch.equal(a, d), "load_state_dict back to the original weights must restore the original result"


def test_invalidate_folded_after_data_write(weights):
    """Writes through `.data` change neither the pointer nor autograd's version counter: the caches cannot
    see them, MVSNet.invalidate_folded() is the documented way to drop the folded / packed copies."""
    from test_gpu_parity import load_model
    from scene_3dreconstruction_mvsnet_b200 import synth
    imgs, proj, dv = synth.make_inputs(B=1, V=3, H=64, W=96, D=16, focal=90.0, interval_scale=8.0, seed=6)
    imgs, proj, dv = imgs.to(DEV), proj.to(DEV), dv.to(DEV)
    m = load_model(weights, precision="bf16")
    with torch.no_grad():
        a = m(imgs, proj, dv)["depth"].clone()
        m.cost_regularization.conv2.conv.weight.data.mul_(1.25)
        m.feature.conv3.conv.weight.data.mul_(0.75)
        m.invalidate_folded()
        b = m(imgs, proj, dv)["depth"].clone()
        fresh = load_model(weights, precision="bf16")
        fresh.cost_regularization.conv2.conv.weight.mul_(1.25)
        fresh.feature.conv3.conv.weight.mul_(0.75)
        c = fresh(imgs, proj, dv)["depth"]
    assert not torch.equal(a, b) and torch.equal(b, c)


def test_cache_clears_from_another_thread_do_not_break_forwards(weights):
    """mvs_weight_cache_clear() takes the cache's lifetime lock exclusively: a forward pass on another host thread either
    finishes enqueueing with its packed weights intact or starts after the clear and re-packs (the old
    two-clear graveyard could free a buffer between a lookup and its launch)."""
    import threading
    from test_gpu_parity import load_model
    from scene_3dreconstruction_mvsnet_b200 import ops, synth
    inp = tuple(t.to(DEV) for t in synth.make_inputs(B=1, V=3, H=64, W=96, D=16, focal=90.0, interval_scale=8.0, seed=13))
    m = load_model(weights, precision="bf16")
    with torch.no_grad():
        want = m(*inp)["depth"].clone()
    torch.cuda.synchronize()
    stop, errors = [False], []

    def clearer():
        while not stop[0]:
            ops.weights_changed()

    t = threading.Thread(target=clearer)
    t.start()
    try:
        with torch.no_grad():
            for _ in range(40):
                got = m(*inp)["depth"]
                if not torch.equal(got, want):
                    errors.append("result changed")
                    break
    finally:
        stop[0] = True
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors


def test_two_host_threads_two_streams_same_results(weights):
    """nn.DataParallel calls forward from one host thread per replica (train.py:125): two threads, each with its own
    model replica and CUDA stream on the same device, must reproduce the serial results bit for bit (thread-local
    plans, per-stream workspaces, the shared packed-weight cache)."""
    import threading
    from test_gpu_parity import load_model
    from scene_3dreconstruction_mvsnet_b200 import synth
    inputs = [tuple(t.to(DEV) for t in synth.make_inputs(B=1, V=3, H=64, W=96, D=16, focal=90.0, interval_scale=8.0, seed=s))
              for s in (11, 12)]
    models = [load_model(weights, precision="bf16") for _ in range(2)]
    with torch.no_grad():
        serial = [m(*inp)["depth"].clone() for m, inp in zip(models, inputs)]
    torch.cuda.synchronize()
    results, errors = [None, None], []

    def work(i):
        try:
            s = torch.cuda.Stream(DEV)
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(8):
                    out = models[i](*inputs[i])["depth"]
                results[i] = out.clone()
            s.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert torch.equal(results[0], serial[0]) and torch.equal(results[1], serial[1])


def test_two_devices_from_one_process(weights):
    """nn.DataParallel-style use (train.py:125): replicas on two devices driven from two host threads of one process.
    Needs two GPUs; skipped otherwise."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import threading
    from test_gpu_parity import load_model
    from scene_3dreconstruction_mvsnet_b200 import synth
    imgs, proj, dv = synth.make_inputs(B=1, V=3, H=64, W=96, D=16, focal=90.0, interval_scale=8.0, seed=21)
    with torch.no_grad():
        ref = load_model(weights, precision="bf16")(imgs.to(DEV), proj.to(DEV), dv.to(DEV))["depth"].cpu()
    results, errors = {}, []

    def work(d):
        try:
            dev = "cuda:%d" % d
            m = load_model(weights, precision="bf16").to(dev)
            with torch.no_grad():
                for _ in range(4):
                    out = m(imgs.to(dev), proj.to(dev), dv.to(dev))["depth"]
            results[d] = out.cpu()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert torch.equal(results[0], ref) and torch.equal(results[1], ref)
