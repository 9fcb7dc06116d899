"""Where the tensor-core mode's error comes from: logits at the golden sample pixels (reference outputs at C1 / C2)
with the stages switched from fp32 to their tensor-core form one at a time.
    python tools/error_budget.py [c1|c2]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_golden
from scene_3dreconstruction_mvsnet_b200 import ops, synth
from scene_3dreconstruction_mvsnet_b200.models import MVSNet

torch.backends.cudnn.allow_tf32 = False
tag = sys.argv[1] if len(sys.argv) > 1 else "c1"
g = load_golden("config_%s.npz" % tag)
w = load_golden("weights_calibrated.npz")
imgs, proj, dv = (t.cuda() for t in synth.make_named(str(g["name"]), seed=0))
m = MVSNet(refine=False, precision="bf16")
m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
m = m.cuda().eval()
m32 = MVSNet(refine=False, precision="fp32")   # its extract_features keeps cuDNN in fp32 (the tensor-core modes allow TF32)
m32.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
m32 = m32.cuda().eval()
yx = g["sample_yx"]
ref = g["logits_samples"]


def report(name, logits):
    lg = logits[0][:, yx[:, 0], yx[:, 1]].float().cpu().numpy()
    e = np.abs(lg - ref)
    print("%-58s logits err max %.3e  mean %.3e  rms %.3e" % (name, e.max(), e.mean(), np.sqrt((e ** 2).mean())), flush=True)


with torch.no_grad():
    fea32 = m32.extract_features(imgs)                     # cuDNN fp32 (TF32 off)
    var32 = ops.warp_variance(fea32, proj, dv)
    folded = m.cost_regularization.folded_params()
    report("fp32 features, fp32 warp, fp32 CostRegNet", ops.cost_regularization(var32, folded, "fp32"))
    report("fp32 features, fp32 warp, tc CostRegNet (fp32 volume in)", ops.cost_regularization(var32, folded, "bf16"))
    report("fp32 features -> fp16 texels, tc warp, tc CostRegNet", ops.warp_variance_costreg_bf16(fea32, proj, dv, folded))
    feat = ops.featurenet_tc(imgs, m.feature.folded_native())
    report("tc features, tc warp, tc CostRegNet (the mode)", ops.warp_variance_costreg_bf16(feat, proj, dv, folded))
    f16 = feat.to_nchw()
    var_t = ops.warp_variance(f16, proj, dv)
    report("tc features, fp32 warp, fp32 CostRegNet", ops.cost_regularization(var_t, folded, "fp32"))
    report("tc features, fp32 warp, tc CostRegNet", ops.cost_regularization(var_t, folded, "bf16"))
    f16r = fea32.half().float()
    var_r = ops.warp_variance(f16r, proj, dv)
    report("fp32 features rounded to fp16, fp32 warp, fp32 CostRegNet", ops.cost_regularization(var_r, folded, "fp32"))
    del var_r
    fe = (f16 - fea32).abs()
    print("features: max |f| %.3f, tc-vs-fp32 err max %.3e mean %.3e" % (fea32.abs().max().item(), fe.max().item(), fe.mean().item()))
    del var_t
    # the volume itself: tc warp on fp32 features (fp16 texels) against the fp32 volume
    vol = ops.warp_variance_cp8(fea32, proj, dv)
    v = vol[0].float().permute(0, 4, 1, 2, 3).reshape(32, *vol.shape[2:5])
    ev = (v - var32[0]).abs()
    print("volume: max var %.3f, tc-warp-vs-fp32 err max %.3e mean %.3e; rel-to-(|ref|+1e-3) max %.3e" % (
        var32.max().item(), ev.max().item(), ev.mean().item(), (ev / (var32[0].abs() + 1e-3)).max().item()))
