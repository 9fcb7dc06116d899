"""Experiment: FeatureNet with BN folded into the convs and cuDNN's fused conv+bias+relu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200.models.mvsnet import FeatureNet
torch.manual_seed(0)
net = FeatureNet().cuda().eval()
for m in net.modules():
    if isinstance(m, torch.nn.BatchNorm2d):
        m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
x = torch.rand(5, 3, 1152, 1600, device="cuda").contiguous(memory_format=torch.channels_last)
torch.backends.cudnn.benchmark = True
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
layers = []
for name in ["conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6"]:
    m = getattr(net, name)
    scale = m.bn.weight / (m.bn.running_var + m.bn.eps).sqrt()
    w = (m.conv.weight * scale.view(-1, 1, 1, 1)).detach().contiguous(memory_format=torch.channels_last)
    b = (m.bn.bias - m.bn.running_mean * scale).detach()
    layers.append((w, b, m.conv.stride, m.conv.padding))
def fused(x):
    for w, b, s, p in layers:
        x = torch.cudnn_convolution_relu(x, w, b, s, p, (1, 1), 1)
    return torch.nn.functional.conv2d(x, net.feature.weight, net.feature.bias, 1, 1)
def fused_lp(x, dt):
    x = x.to(dt)
    for w, b, s_, p in layers:
        x = torch.cudnn_convolution_relu(x, w.to(dt), b.to(dt), s_, p, (1, 1), 1)
    return torch.nn.functional.conv2d(x, net.feature.weight.to(dt), net.feature.bias.to(dt), 1, 1)
with torch.no_grad():
    netc = net.to(memory_format=torch.channels_last)
    ref = netc(x)
    for dt in (torch.float16, torch.bfloat16):
        ws = [(w.to(dt), b.to(dt), s_, p) for w, b, s_, p in layers]
        fw, fb = net.feature.weight.to(dt).contiguous(memory_format=torch.channels_last), net.feature.bias.to(dt)
        def run():
            y = x.to(dt)
            for w, b, s_, p in ws:
                y = torch.cudnn_convolution_relu(y, w, b, s_, p, (1, 1), 1)
            return torch.nn.functional.conv2d(y, fw, fb, 1, 1)
        y = run()
        print(dt, "%.3f ms" % timeit(run), "maxdiff %.3g (absmax %.3g)" % ((y.float() - ref).abs().max().item(), ref.abs().max().item()))
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        netc = net.to(memory_format=torch.channels_last)
        ref = netc(x)
        t0 = timeit(lambda: netc(x))
        y = fused(x)
        t1 = timeit(lambda: fused(x))
        print("tf32=%s  module %.3f ms   folded+fused %.3f ms   maxdiff %.3g (absmax %.3g)" % (tf32, t0, t1, (y - ref).abs().max().item(), ref.abs().max().item()))
