"""Throughput of back-to-back forwards on one stream against the same forwards alternating over two streams (independent
depth maps: the tail of one kernel overlaps the next map's kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scene_3dreconstruction_mvsnet_b200 import synth
from scene_3dreconstruction_mvsnet_b200.models import MVSNet

name = sys.argv[1] if len(sys.argv) > 1 else "c2_dtu_5view_1152x1600"
n = 40
torch.manual_seed(1)
model = MVSNet(refine=False, precision="bf16").cuda().eval()
inp = [t.cuda() for t in synth.make_named(name)]
streams = [torch.cuda.Stream() for _ in range(3)]
with torch.no_grad():
    for ns in (1, 2, 3, 1, 2):
        for s in streams[:ns]:
            with torch.cuda.stream(s):
                for _ in range(3):
                    model(*inp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams[:ns]:
            s.wait_event(e0)
        for i in range(n):
            with torch.cuda.stream(streams[i % ns]):
                model(*inp)
        for s in streams[:ns]:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        print("%s: %d stream(s): %.3f ms per depth map = %.1f depth maps/s" % (name, ns, e0.elapsed_time(e1) / n, n * 1e3 / e0.elapsed_time(e1)), flush=True)
