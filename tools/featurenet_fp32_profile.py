"""Times the strict-fp32 FeatureNet (csrc/conv2d_fp32.cu) at the DTU shape, whole and per layer, next to the cuDNN path it
replaces in precision='fp32' inference.   python tools/featurenet_fp32_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import ops
from scene_3dreconstruction_mvsnet_b200.models.mvsnet import FeatureNet

torch.manual_seed(0)
net = FeatureNet().cuda().eval()
N, H, W = 5, 1152, 1600
x = torch.rand(1, N, 3, H, W, device="cuda")

def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

with torch.no_grad():
    folded = net.folded_native()
    fea = ops.featurenet_fp32(x, folded)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        xc = x[0].contiguous(memory_format=torch.channels_last)
        ref = net.infer(xc)
        t_cudnn = timeit(lambda: net.infer(xc))
    print("cuDNN fp32 (NHWC, fused conv+bias+ReLU, TF32 off): %.3f ms" % t_cudnn)
    print("own fp32 kernels: %.3f ms   max |diff| vs cuDNN %.3g (feature absmax %.3g)" % (
        timeit(lambda: ops.featurenet_fp32(x, folded)), (fea[0] - ref).abs().max().item(), ref.abs().max().item()))
    cin = [3, 8, 8, 16, 16, 16, 32, 32]
    h, w = H, W
    tot = 0.0
    for l, (wt, sh) in enumerate(folded):
        k = wt.shape[2]
        s = 2 if k == 5 else 1
        a = torch.randn(N, cin[l], h, w, device="cuda")
        ms = timeit(lambda: ops.conv2d_bn_relu(a, wt, sh, relu=l != 7, stride=s))
        h, w = (h - 1) // s + 1, (w - 1) // s + 1
        flop = 2.0 * N * wt.shape[0] * cin[l] * k * k * h * w
        tot += ms
        print("layer %d %2d->%2d k%d s%d -> %4dx%4d: %.3f ms = %.1f TFLOP/s" % (l, cin[l], wt.shape[0], k, s, h, w, ms, flop / ms / 1e9))
    print("sum of layers %.3f ms" % tot)
