import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from scene_3dreconstruction_mvsnet_b200 import synth
from scene_3dreconstruction_mvsnet_b200.models import MVSNet, mvsnet_loss
dev = "cuda"
torch.manual_seed(1)
model = MVSNet(refine=False).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
B = 4
imgs, proj, dv = synth.make_inputs(B=B, V=3, H=512, W=640, D=192, focal=361.5, interval_scale=1.06, seed=0)
imgs, proj, dv = imgs.to(dev), proj.to(dev), dv.to(dev)
gt = torch.full((B, 128, 160), 650.0, device=dev); mask = torch.ones_like(gt)
def step():
    opt.zero_grad(set_to_none=True)
    out = model(imgs, proj, dv)
    mvsnet_loss(out["depth"], gt, mask).backward()
    opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
