"""Depth / confidence error of the three precision modes against the reference goldens and against the strict mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_golden
from scene_3dreconstruction_mvsnet_b200 import synth
from scene_3dreconstruction_mvsnet_b200.models import MVSNet
torch.backends.cudnn.allow_tf32 = False
w = load_golden("weights_calibrated.npz")
def model(p):
    m = MVSNet(refine=False, precision=p)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return m.cuda().eval()
ms = {p: model(p) for p in ("fp32", "bf16", "fast")}
for case in ("case_a.npz", "case_b.npz"):
    c = load_golden(case)
    rng = float(c["dv"].max() - c["dv"].min())
    for p, m in ms.items():
        with torch.no_grad():
            o = m(torch.from_numpy(c["imgs"]).cuda(), torch.from_numpy(c["proj"]).cuda(), torch.from_numpy(c["dv"]).cuda())
        e = np.abs(o["depth"].cpu().numpy() - c["depth"])
        ce = np.abs(o["photometric_confidence"].cpu().numpy() - c["conf"])
        print("%s %-5s depth err / range: max %.2e mean %.2e | conf abs err: max %.2e mean %.2e" % (case[:6], p, e.max() / rng, e.mean() / rng, ce.max(), ce.mean()))
del ms["fast"]
imgs, proj, dv = synth.make_named("c1_3view_512x640")
imgs, proj, dv = imgs.cuda(), proj.cuda(), dv.cuda()
rng = float(dv.max() - dv.min())
with torch.no_grad():
    ref = ms["fp32"](imgs, proj, dv)
    for p in ("bf16",):
        o = ms[p](imgs, proj, dv)
        e = (o["depth"] - ref["depth"]).abs()
        ce = (o["photometric_confidence"] - ref["photometric_confidence"]).abs()
        print("C1 full size, %-5s vs fp32 mode: depth err / range max %.2e mean %.2e | conf abs err max %.2e mean %.2e" % (p, e.max().item() / rng, e.mean().item() / rng, ce.max().item(), ce.mean().item()))

# ---- BASELINE.json shapes against the unmodified reference's outputs (tests/golden/config_*.npz)
import json
from test_gpu_config_goldens import TAGS, measure
print("\nconfig-sized goldens (reference outputs at C1 / C3 / C2):")
for tag in TAGS:
    for p in ("fp32", "bf16", "reference_cuda_default"):
        torch.cuda.empty_cache()
        r = measure(tag, w, p)
        print(tag, p, json.dumps({k: float("%.3g" % v) for k, v in r.items()}), flush=True)
