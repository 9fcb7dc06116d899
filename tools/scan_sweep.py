"""BASELINE config 5 in miniature: one DTU-shaped scan (49 reference views, 5 views each at 1152x1600, D=192) swept
through the whole pipeline -- 8-bit host images -> DepthMapRunner (H2D, FeatureNet, warp+variance, CostRegNet, tail, D2H)
sharded over the ranks by reference view -> gather of the depth / confidence maps on rank 0 -> geometric-consistency
filter of every view against its 10 neighbours (fusion.filter_view).  Synthetic images / cameras, random-init weights.
    python tools/scan_sweep.py                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/scan_sweep.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from scene_3dreconstruction_mvsnet_b200 import fusion, sharding, synth
from scene_3dreconstruction_mvsnet_b200.models import MVSNet
from scene_3dreconstruction_mvsnet_b200.runner import DepthMapRunner

NV, V, H, W, D, NFILTER = int(os.environ.get("SWEEP_VIEWS", 49)), 5, 1152, 1600, 192, 10
rank, world = sharding.rank_world()
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

# cameras on an arc (feature-resolution intrinsics like the reference's cam files scaled by 1/4), 8-bit images
h, w = H // 4, W // 4
K = np.array([[723.0, 0, w / 2.0], [0, 723.0, h / 2.0], [0, 0, 1]], np.float64)
Es = []
for i in range(NV):
    a = 0.01 * (i - NV / 2.0)
    E = np.eye(4)
    E[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
    E[:3, 3] = [-25.0 * (i - NV / 2.0), 4.0 * (i % 3), 0.0]
    Es.append(E)
proj = []
for E in Es:
    P = E.copy()
    P[:3, :4] = K @ E[:3, :4]
    proj.append(P.astype(np.float32))
g = torch.Generator().manual_seed(0)
images = [torch.randint(0, 256, (3, H, W), dtype=torch.uint8, generator=g) for _ in range(NV)]
neigh = lambda i, n: [j for j in sorted(range(NV), key=lambda j: (abs(j - i), j)) if j != i][:n]  # pair.txt stand-in
dv = (425.0 + 2.65 * torch.arange(D, dtype=torch.float32)).unsqueeze(0)

torch.manual_seed(1)
runner = DepthMapRunner(MVSNet(refine=False, precision="bf16"), device=str(dev))
mine = sharding.shard_indices(NV, rank, world)

# what a loader with worker processes would hand over: assembled, page-locked inputs per owned reference view
batches = []
for i in mine:
    ids = [i] + neigh(i, V - 1)
    batches.append((torch.stack([images[j] for j in ids]).unsqueeze(0).pin_memory(),
                    torch.from_numpy(np.stack([proj[j] for j in ids])).unsqueeze(0).pin_memory(), dv.pin_memory()))

def views():
    return iter(batches)

runner.run_views(batches[:2])  # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = runner.run_views(views())
torch.cuda.synchronize()
t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
t_depth = float(t.item())
maps = sharding.gather_maps(mine, [torch.from_numpy(np.stack([d[0], c[0]])) for d, c in res], NV)
if rank == 0:
    depth = np.stack([m[0].numpy() for m in maps]); conf = np.stack([m[1].numpy() for m in maps])
    nb0 = neigh(0, NFILTER)
    for _ in range(3):  # warm-up at the real size
        fusion.filter_view(depth[0], conf[0], K, Es[0], depth[nb0], [K] * len(nb0), [Es[j] for j in nb0])
    t0 = time.perf_counter()
    kept = 0.0
    for i in range(NV):
        nb = neigh(i, NFILTER)
        out = fusion.filter_view(depth[i], conf[i], K, Es[i], depth[nb], [K] * len(nb), [Es[j] for j in nb])
        kept += out["final_mask"].mean()
    t_fuse = time.perf_counter() - t0
    print(json.dumps({"config": "C5 miniature: 1 scan, %d reference views x %d views %dx%d, D=%d" % (NV, V, H, W, D), "n_gpus": world,
                      "depth_maps_per_s_e2e_uint8": NV / t_depth, "depth_sweep_s": t_depth,
                      "fusion_views_per_s": NV / t_fuse, "fusion_s": t_fuse, "fusion_src_views": NFILTER,
                      "mean_final_mask": kept / NV}))
if world > 1:
    dist.destroy_process_group()
