"""Full-size (C2) comparison of the TMA-window kernel with the strict fp32 kernel: error statistics and where violations sit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import ops, synth
B, V, h, w, D = 1, 5, 288, 400, 192
fea = synth.make_features(B, V, 32, h, w, seed=4).half().float().cuda()
_, proj, dv = synth.make_named("c2_dtu_5view_1152x1600")
proj, dv = proj.cuda(), dv.cuda()
cp8 = ops.warp_variance_cp8(fea, proj, dv)
ref = ops.warp_variance(fea, proj, dv)
got = cp8.permute(0, 1, 5, 2, 3, 4).reshape(B, 32, D, h, w).float()
err = (got - ref).abs()
tol = ref.abs() * 2.0 ** -7 + 8e-3
bad = err > tol
print("hacc=%s elements %d violations %d max err %.4g max err/tol %.3f mean err %.3g" % (
    os.environ.get("MVS_WIN_HACC", "default"), err.numel(), int(bad.sum()), err.max().item(), (err / tol).max().item(), err.mean().item()))
if bad.any():
    idx = bad.nonzero()
    print("first violations (b,c,d,y,x):", idx[:8].tolist())
    print("by d:", torch.bincount(idx[:, 2], minlength=D).nonzero().flatten()[:20].tolist())
    print("x range", int(idx[:, 4].min()), int(idx[:, 4].max()), "y range", int(idx[:, 3].min()), int(idx[:, 3].max()))
    i = idx[0]
    print("example: got %.5f ref %.5f" % (got[tuple(i)].item(), ref[tuple(i)].item()))
# relative to a bf16 ulp of the reference
ulp = ref.abs().clamp_min(1e-3) * 2.0 ** -8
print("err in bf16 ulps: max %.2f, p99.99 %.2f" % ((err / ulp).max().item(), torch.quantile((err / ulp).flatten()[::64], 0.9999).item()))
