"""Host-side cost of enqueueing one forward pass (plans, tensor maps, launches) against its GPU time: the GPU must stay
the bound.  Enqueues N forwards back to back and reports host time per forward (before the final synchronize) and
device time per forward."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import synth
from scene_3dreconstruction_mvsnet_b200.models import MVSNet

name = sys.argv[1] if len(sys.argv) > 1 else "c2_dtu_5view_1152x1600"
n = 20
torch.manual_seed(1)
model = MVSNet(refine=False, precision="bf16").cuda().eval()
inp = [t.cuda() for t in synth.make_named(name)]
with torch.no_grad():
    for _ in range(5):
        model(*inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        model(*inp)
    e1.record()
    host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
print("%s: host enqueue %.3f ms per forward, device %.3f ms per forward" % (name, host * 1e3, e0.elapsed_time(e1) / n))
