"""Times the fused warp+variance kernel alone at a named workload's shape (rectified or rotated cameras); the command
profiled under ncu for the warp kernel.
    python tools/warp_profile.py [workload] [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import ops, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c2_dtu_5view_1152x1600"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
V, H, W, D, _, _ = synth.config_of(name)
fea = synth.make_features(1, V, 32, H // 4, W // 4, seed=0).cuda()
t16 = fea.half().view(V, 4, 8, H // 4, W // 4).permute(0, 3, 1, 4, 2).contiguous()
rcp8 = ops.Rcp8Features(t16, 1, V, H // 4, W // 4)
_, proj, dv = synth.make_named(name)
proj, dv = proj.cuda(), dv.cuda()
for _ in range(3):
    ops.warp_variance_cp8(rcp8, proj, dv)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    ops.warp_variance_cp8(rcp8, proj, dv)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
alg = 2 * 32 * D * (H // 4) * (W // 4) + 4 * V * 32 * (H // 4) * (W // 4)
print("%s: %.3f ms per call (incl. homography compose + output allocation), %.0f GB/s algorithmic" % (name, ms, alg / ms / 1e6), flush=True)
