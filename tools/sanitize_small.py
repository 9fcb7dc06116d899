"""Small end-to-end run of every tensor-core-mode kernel (for compute-sanitizer --tool memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import synth, ops
from scene_3dreconstruction_mvsnet_b200.models import MVSNet

torch.manual_seed(1)
m = MVSNet(refine=False, precision="bf16").cuda().eval()
for (V, H, W, D, yaw) in ((3, 64, 96, 16, 0.03), (5, 96, 160, 24, 0.2), (2, 32, 96, 8, 0.0)):
    imgs, proj, dv = synth.make_inputs(B=1, V=V, H=H, W=W, D=D, focal=0.9 * W / 4, interval_scale=8.0, yaw=yaw, seed=V)
    with torch.no_grad():
        out = m(imgs.cuda(), proj.cuda(), dv.cuda())
        out8 = m((imgs * 255).round().to(torch.uint8).cuda(), proj.cuda(), dv.cuda())
    torch.cuda.synchronize()
    print(V, H, W, D, float(out["depth"].mean()), float(out8["photometric_confidence"].mean()))
# stress geometry for the window planner: strong zoom -> global-gather fallback
fea = synth.make_features(1, 3, 32, 24, 72, seed=11).cuda()
proj = torch.eye(4).repeat(1, 3, 1, 1)
proj[0, :, 0, 0] = proj[0, :, 1, 1] = 60.0
proj[0, 1, 0, 0] = proj[0, 1, 1, 1] = 180.0
proj[0, :, 0, 2], proj[0, :, 1, 2] = 36.0, 12.0
proj[0, 1, 0, 3], proj[0, 2, 0, 3] = 40.0 * 180, -40.0 * 60
dv = (425 + 20.0 * torch.arange(12, dtype=torch.float32)).unsqueeze(0)
v = ops.warp_variance_cp8(fea, proj.cuda(), dv.cuda())
torch.cuda.synchronize()
print("stress", float(v.float().abs().mean()))
