"""Per-kernel device times of one MVSNet forward at the DTU shape (default precision mode), in launch order, from
CUPTI via torch.profiler (back-to-back launches, warm caches -- unlike the serialised ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from scene_3dreconstruction_mvsnet_b200 import synth
from scene_3dreconstruction_mvsnet_b200.models import MVSNet

torch.manual_seed(1)
model = MVSNet(refine=False, precision=sys.argv[1] if len(sys.argv) > 1 else "bf16").cuda().eval()
imgs, proj, dv = (t.cuda() for t in synth.make_named("c2_dtu_5view_1152x1600"))
with torch.no_grad():
    for _ in range(5):
        model(imgs, proj, dv)
    torch.cuda.synchronize()
    n = 5
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            model(imgs, proj, dv)
        torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
per = len(evs) // n
tot = 0.0
for i in range(per):
    us = sum(evs[k * per + i].time_range.elapsed_us() for k in range(n)) / n
    tot += us
    print("%3d %8.1f us  %s" % (i, us, evs[i].name[:90]))
print("sum of kernels %.1f us; span %.1f us" % (tot, (evs[-1].time_range.end - evs[-per].time_range.start)))
