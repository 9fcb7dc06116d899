"""Times the (out-of-scope) cuDNN FeatureNet at the DTU shape under different precision settings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200.models.mvsnet import FeatureNet

torch.manual_seed(0)
net = FeatureNet().cuda().eval()
x = torch.rand(5, 3, 1152, 1600, device="cuda")
torch.backends.cudnn.benchmark = True

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

with torch.no_grad():
    torch.backends.cudnn.allow_tf32 = False
    ref = net(x)
    print("fp32 strict      %.3f ms" % timeit(lambda: net(x)))
    torch.backends.cudnn.allow_tf32 = True
    y = net(x)
    print("tf32 allowed     %.3f ms  maxerr %.3g (ref absmax %.3g)" % (timeit(lambda: net(x)), (y - ref).abs().max().item(), ref.abs().max().item()))
    xc = x.contiguous(memory_format=torch.channels_last)
    netc = net.to(memory_format=torch.channels_last)
    y = netc(xc)
    print("tf32 + NHWC      %.3f ms  maxerr %.3g" % (timeit(lambda: netc(xc)), (y - ref).abs().max().item()))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = netc(xc)
        print("bf16 autocast NHWC %.3f ms  maxerr %.3g" % (timeit(lambda: netc(xc)), (y.float() - ref).abs().max().item()))
    with torch.autocast("cuda", dtype=torch.float16):
        y = netc(xc)
        print("fp16 autocast NHWC %.3f ms  maxerr %.3g" % (timeit(lambda: netc(xc)), (y.float() - ref).abs().max().item()))
