"""Geometric-consistency filter at the DTU depth-map size (288x400, 10 source views): the CUDA kernel (device-resident
inputs, and through the numpy-in / numpy-out mirror) next to the numpy + cv2-restatement port on the host cores."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import fusion_oracle as fo
from scene_3dreconstruction_mvsnet_b200 import fusion, _lib

rs = np.random.RandomState(0)
h, w, S = 288, 400, 10
K = np.array([[0.9 * w, 0, w / 2.0], [0, 0.9 * w, h / 2.0], [0, 0, 1]], np.float64)
ref = (600 + 30 * rs.rand(h, w)).astype(np.float32)
conf = rs.rand(h, w).astype(np.float32)
Ks, Es, Ds = [], [], []
for s in range(S):
    E = np.eye(4); E[:3, 3] = [30.0 * (s - 5), 5.0 * s, 2.0 * s]
    Ks.append(K); Es.append(E); Ds.append((600 + 30 * rs.rand(h, w)).astype(np.float32))
Ds, Ks, Es = np.stack(Ds), np.stack(Ks), np.stack(Es)
t0 = time.perf_counter(); fo.filter_view(ref, conf, K, np.eye(4), Ds, Ks, Es); t_cpu = time.perf_counter() - t0
fusion.filter_view(ref, conf, K, np.eye(4), Ds, Ks, Es)
t0 = time.perf_counter()
for _ in range(10): fusion.filter_view(ref, conf, K, np.eye(4), Ds, Ks, Es)
t_api = (time.perf_counter() - t0) / 10
# kernel only
lib = _lib.load(); dev = torch.device("cuda:0")
d_ref, d_conf, d_src = torch.as_tensor(ref).to(dev), torch.as_tensor(conf).to(dev), torch.as_tensor(Ds).to(dev)
avg = torch.empty((h, w), dtype=torch.float64, device=dev); gs = torch.empty((h, w), dtype=torch.int32, device=dev)
pm, gm, fm = (torch.empty((h, w), dtype=torch.uint8, device=dev) for _ in range(3))
p = lambda t: ctypes.c_void_p(t.data_ptr()); hp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
E0 = np.eye(4)
def run():
    lib.mvs_filter_depth(p(d_ref), p(d_conf), hp(K), hp(E0), p(d_src), hp(Ks), hp(Es), S, h, w, 1.0, 0.01, 3, 0.8, p(avg), p(gs), p(pm),
                         p(gm), p(fm), None, None, None, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); torch.cuda.synchronize()
t_k = e0.elapsed_time(e1) / 50
byts = (S + 2) * h * w * 4 + h * w * (8 + 4 + 3)
print("filter_view 288x400, %d source views: numpy port %.1f ms on %d threads | CUDA, host arrays in/out %.2f ms | kernel + camera upload %.3f ms "
      "(%.1f MB algorithmic -> %.0f GB/s)" % (S, t_cpu * 1e3, os.cpu_count(), t_api * 1e3, t_k, byts / 1e6, byts / t_k / 1e6))
