"""Per-role cycle breakdown of the FeatureNet layers on the tensor-core kernel at the DTU shape (5 x 1152x1600)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import ops, _lib
lib = _lib.load()
N, H, W = 5, 1152, 1600
layers = [("conv0", 3, 8, 3, H, W, False), ("conv1", 8, 8, 3, H, W, True), ("conv2", 8, 16, 5, H, W, False),
          ("conv3", 16, 16, 3, H // 2, W // 2, False), ("conv4", 16, 16, 3, H // 2, W // 2, True),
          ("conv5", 16, 32, 5, H // 2, W // 2, False), ("conv6", 32, 32, 3, H // 4, W // 4, False)]
tot = 0.0
for name, ci, co, k, h, w, s2d in layers:
    x = torch.randn(N, ci, h, w, device="cuda")
    wt = torch.randn(co, ci, k, k, device="cuda") * 0.05
    sh = torch.zeros(co, device="cuda")
    dbg = torch.zeros(148 * 12, dtype=torch.int64, device="cuda")
    run = lambda: ops.conv2d_bn_relu_tc(x, wt, sh, relu=True, stride=2 if k == 5 else 1, s2d_out=s2d)
    run(); torch.cuda.synchronize()
    lib.mvs_tc_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
    run(); torch.cuda.synchronize()
    lib.mvs_tc_set_debug_buffer(None)
    t = dbg.view(148, 12).double()
    t = t[t[:, 5] > 0]
    steps = t[:, 5].mean().item()
    m = t.mean(0) / steps
    ms = t[:, 4].max().item() / 1.965e6
    tot += ms
    print("%-6s %2d->%2d k%d %4dx%4d steps/CTA %6.1f (%.3f ms) | per step: total %6.0f  mma: wait_full %5.0f wait_tmem %5.0f issue %5.0f | epi: wait %5.0f work %5.0f | producer wait_empty %5.0f | release %5.0f"
          % (name, ci, co, k, h, w, steps, ms, m[4], m[1], m[2], m[3], m[6], m[7], m[0], m[8]))
    del x
print("sum of tensor-core layer times (without feature layer = conv6 shape again): %.3f ms" % tot)
