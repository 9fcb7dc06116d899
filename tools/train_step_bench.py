"""BASELINE config 4: training step (forward + backward + Adam) at 3 views 512x640, D=192, batch 4 per GPU.
Forward/backward of the fused warp+variance op are ours; FeatureNet / CostRegNet / softmax run on cuDNN + autograd
(train-mode BatchNorm needs batch statistics).  Optional DDP over NCCL when launched with torchrun."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from scene_3dreconstruction_mvsnet_b200 import synth, _lib
from scene_3dreconstruction_mvsnet_b200.models import MVSNet, mvsnet_loss

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B = int(os.environ.get("TRAIN_B", 4))
steps, warmup = 5, 2
torch.manual_seed(1)
model = MVSNet(refine=False).to(dev).train()
if world > 1:
    model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
imgs, proj, dv = synth.make_inputs(B=B, V=3, H=512, W=640, D=192, focal=361.5, interval_scale=1.06, seed=rank)
imgs, proj, dv = imgs.to(dev), proj.to(dev), dv.to(dev)
gt = torch.full((B, 128, 160), 650.0, device=dev)
mask = torch.ones_like(gt)
ev = []
def step(record=False):
    opt.zero_grad(set_to_none=True)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    out = model(imgs, proj, dv)
    loss = mvsnet_loss(out["depth"], gt, mask)
    e[1].record()
    loss.backward()
    opt.step()
    e[2].record()
    if record: ev.append(e)
    return loss
for _ in range(warmup): step()
torch.cuda.synchronize()
n0 = _lib.launch_count()
t0 = time.perf_counter()
for _ in range(steps): l = step(True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
fwd = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
bwd = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
if rank == 0:
    print(json.dumps({"config": "C4 train step, B=%d/GPU, V=3, 512x640, D=192" % B, "n_gpus": world, "ms_per_step": dt * 1e3,
                      "fwd_ms": fwd, "bwd_opt_ms": bwd, "samples_per_s": world * B / dt, "loss": float(l),
                      "our_launches_per_step": (_lib.launch_count() - n0) / steps,
                      "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
if world > 1:
    dist.destroy_process_group()
