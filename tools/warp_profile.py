"""Times one of the fused warp+variance kernels alone at a named workload's shape (rectified or rotated cameras); the
command profiled under ncu for these kernels.
    python tools/warp_profile.py [workload] [iterations] [tc|fp32|bwd] [batch]
  tc    warp_variance_win_kernel (fp16 texels, fp16 CP8 volume; the bench default)
  fp32  warp_variance_fwd2_kernel (strict fp32, the reference's precision)
  bwd   warp_volume_bwd_kernel (backward of the fused op, training; use workload c1_3view_512x640 and batch 4 for C4)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import ops, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c2_dtu_5view_1152x1600"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
mode = sys.argv[3] if len(sys.argv) > 3 else "tc"
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
V, H, W, D, _, _ = synth.config_of(name)
h, w = H // 4, W // 4
fea = synth.make_features(B, V, 32, h, w, seed=0).cuda()
_, proj, dv = synth.make_named(name, B=B)
proj, dv = proj.cuda(), dv.cuda()
if mode == "tc":
    t16 = fea.half().view(B * V, 4, 8, h, w).permute(0, 3, 1, 4, 2).contiguous()
    rcp8 = ops.Rcp8Features(t16, B, V, h, w)
    run = lambda: ops.warp_variance_cp8(rcp8, proj, dv)
    alg = 2 * B * 32 * D * h * w + 4 * B * V * 32 * h * w
elif mode == "fp32":
    pc = ops.compose_like_reference(proj)
    run = lambda: ops.warp_variance_fwd(fea, pc, dv)
    alg = 4 * B * 32 * D * h * w + 4 * B * V * 32 * h * w
else:
    pc = ops.compose_like_reference(proj)
    g = torch.randn(B, 32, D, h, w, device="cuda")
    run = lambda: ops.warp_variance_bwd(g, fea, pc, dv)
    alg = 4 * B * 32 * D * h * w + 3 * 4 * B * V * 32 * h * w
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("%s %s B=%d: %.3f ms per call (incl. homography compose + output allocation), %.0f GB/s algorithmic"
      % (name, mode, B, ms, alg / ms / 1e6), flush=True)
