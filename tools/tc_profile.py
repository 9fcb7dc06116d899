"""Per-role cycle breakdown of one tensor-core layer (cycle counters written by the kernel itself)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import ops, _lib
lib = _lib.load()
D, H, W = 192, 288, 400
layers = [("conv0", 0, 32, 8, D, H, W), ("conv2", 0, 16, 16, D // 2, H // 2, W // 2), ("conv11", 2, 16, 8, D // 2, H // 2, W // 2),
          ("prob", 0, 8, 1, D, H, W), ("conv1", 1, 8, 16, D, H, W), ("conv9", 2, 32, 16, D // 4, H // 4, W // 4),
          ("conv3", 1, 16, 32, D // 2, H // 2, W // 2), ("conv4", 0, 32, 32, D // 4, H // 4, W // 4)]
for name, kind, ci, co, d, h, w in layers:
    x = torch.randn(1, ci, d, h, w, device="cuda")
    wt = torch.randn((ci, co, 3, 3, 3) if kind == 2 else (co, ci, 3, 3, 3), device="cuda") * 0.05
    sh = torch.zeros(co, device="cuda")
    dbg = torch.zeros(148 * 12, dtype=torch.int64, device="cuda")
    skip = torch.randn(1, co, 2 * d, 2 * h, 2 * w, device="cuda") if kind == 2 and "noskip" not in sys.argv else None
    def run():
        if kind == 2:
            return ops.conv_transpose3d_bn_relu(x, wt, sh, skip=skip, tensor_cores=True)
        return ops.conv3d_bn_relu(x, wt, sh, relu=co != 1, stride=2 if kind == 1 else 1, tensor_cores=True)
    run(); torch.cuda.synchronize()
    lib.mvs_tc_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
    run(); torch.cuda.synchronize()
    lib.mvs_tc_set_debug_buffer(None)
    t = dbg.view(148, 12).double()
    t = t[t[:, 5] > 0]
    steps = t[:, 5].mean().item()
    m = t.mean(0) / steps
    print("%-7s steps/CTA %6.1f | per step cycles: total %7.0f  mma: wait_full %6.0f wait_tmem %6.0f issue %6.0f | epi: wait %6.0f work %6.0f | producer wait_empty %6.0f | release %5.0f"
          % (name, steps, m[4], m[1], m[2], m[3], m[6], m[7], m[0], m[8]))
