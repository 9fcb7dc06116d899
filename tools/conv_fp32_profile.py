"""Times the strict-fp32 CUDA-core convolution layers of CostRegNet at the C2 shape (the command profiled under ncu).
    python tools/conv_fp32_profile.py [layer]     layer: conv0 ... conv7 | conv9 | conv11 | prob | all"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import ops
D, H, W = 192, 288, 400
layers = {"conv0": (0, 32, 8, D, H, W), "conv1": (1, 8, 16, D, H, W), "conv2": (0, 16, 16, D // 2, H // 2, W // 2),
          "conv3": (1, 16, 32, D // 2, H // 2, W // 2), "conv4": (0, 32, 32, D // 4, H // 4, W // 4),
          "conv5": (1, 32, 64, D // 4, H // 4, W // 4), "conv6": (0, 64, 64, D // 8, H // 8, W // 8),
          "conv7": (2, 64, 32, D // 8, H // 8, W // 8), "conv9": (2, 32, 16, D // 4, H // 4, W // 4),
          "conv11": (2, 16, 8, D // 2, H // 2, W // 2), "prob": (0, 8, 1, D, H, W)}
which = sys.argv[1] if len(sys.argv) > 1 else "all"
tot = 0.0
for name, (kind, ci, co, d, h, w) in layers.items():
    if which != "all" and which != name:
        continue
    x = torch.randn(1, ci, d, h, w, device="cuda")
    wt = torch.randn((ci, co, 3, 3, 3) if kind == 2 else (co, ci, 3, 3, 3), device="cuda") * 0.05
    sh = torch.zeros(co, device="cuda")
    run = (lambda: ops.conv_transpose3d_bn_relu(x, wt, sh)) if kind == 2 else (lambda: ops.conv3d_bn_relu(x, wt, sh, relu=co != 1, stride=2 if kind == 1 else 1))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    do, ho, wo = (d // 2, h // 2, w // 2) if kind == 1 else ((2 * d, 2 * h, 2 * w) if kind == 2 else (d, h, w))
    flop = 2.0 * 27 * ci * co * do * ho * wo / (8 if kind == 2 else 1)
    tot += ms
    print("%-6s %2d->%2d %3dx%3dx%3d: %.3f ms = %.1f TFLOP/s" % (name, ci, co, d, h, w, ms, flop / ms / 1e9), flush=True)
    del x
print("sum %.2f ms" % tot)
