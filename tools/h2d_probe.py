"""Aggregate host->device bandwidth with all ranks copying at once: default pinned memory against write-combined
pinned memory (cudaHostAllocWriteCombined), 110.6 MB per copy (the float32 images of one C2 depth map).
    torchrun --nproc-per-node N tools/h2d_probe.py"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 5 * 3 * 1152 * 1600
src = torch.rand(n)
cudart = ctypes.CDLL("libcudart.so.12")
def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_float * (nbytes // 4)).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.float32), p
bufs = {"pinned (torch)": src.pin_memory(), "cudaHostAlloc default": host_alloc(n * 4, 0)[0], "write-combined": host_alloc(n * 4, 4)[0]}
for k in list(bufs)[1:]:
    bufs[k].copy_(src)
dst = torch.empty(n, device=dev)
st = torch.cuda.Stream()
for name, b in bufs.items():
    for _ in range(3):
        dst.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(40):
        dst.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([40 * n * 4 / dt / 1e9], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(gbs)
    if rank == 0:
        print("%d rank(s), %-22s: %.1f GB/s aggregate H2D (%.1f per GPU)" % (world, name, gbs.item(), gbs.item() / world), flush=True)
    assert torch.equal(dst.cpu(), src)
if world > 1:
    dist.destroy_process_group()
