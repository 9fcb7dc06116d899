"""Runs a few MVSNet forwards at the DTU shape (default precision mode) -- the command profiled under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import synth
from scene_3dreconstruction_mvsnet_b200.models import MVSNet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(1)
model = MVSNet(refine=False, precision="bf16").cuda().eval()
imgs, proj, dv = (t.cuda() for t in synth.make_named("c2_dtu_5view_1152x1600"))
with torch.no_grad():
    for _ in range(n):
        out = model(imgs, proj, dv)
torch.cuda.synchronize()
print("ok", float(out["depth"].mean()))
