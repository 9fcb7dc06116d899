"""bench.py workloads beyond the single depth map: BASELINE.json configs[4] (a DTU scan sweep, `c5_scan`) and
configs[3] (the training step, `c4_train`).  Same JSON line contract as bench.py's default workload; `bench` is the
bench.py module (clock sampler, config object, peaks)."""
import json
import os
import time

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------------------------
# c5_scan: one DTU-shaped scan per step -- 49 images 1152x1600, 49 reference views x 5 views, D = 192
# (reference: eval.py:325-396 over datasets/dataloader_eval.py's metas; lists/dtu/test.txt has 22 such scans)
# ------------------------------------------------------------------------------------------------------------------
NV, V_SCAN, H_SCAN, W_SCAN, D_SCAN = 49, 5, 1152, 1600, 192


def make_scan(seed, n_images=NV, H=H_SCAN, W=W_SCAN, D=D_SCAN, nviews=V_SCAN, dtype=torch.float32):
    """Synthetic scan: images on an arc of cameras (feature-resolution intrinsics, like the reference's cam files
    scaled by 1/4), pair list = the nviews-1 nearest cameras (pair.txt stand-in), shared depth hypotheses."""
    h, w = H // 4, W // 4
    K = np.array([[723.0, 0, w / 2.0], [0, 723.0, h / 2.0], [0, 0, 1]], np.float64)
    projs = []
    for i in range(n_images):
        a = 0.01 * (i - n_images / 2.0)
        E = np.eye(4)
        E[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        E[:3, 3] = [-25.0 * (i - n_images / 2.0), 4.0 * (i % 3), 0.0]
        P = E.copy()
        P[:3, :4] = K @ E[:3, :4]
        projs.append(P.astype(np.float32))
    g = torch.Generator().manual_seed(seed)
    if dtype == torch.uint8:
        images = torch.randint(0, 256, (n_images, 3, H, W), dtype=torch.uint8, generator=g)
    else:
        images = torch.rand(n_images, 3, H, W, generator=g)
    pairs = [(i, [j for j in sorted(range(n_images), key=lambda j: (abs(j - i), j)) if j != i][:nviews - 1])
             for i in range(n_images)]
    dv = 425.0 + 2.5 * 1.06 * torch.arange(D, dtype=torch.float32)
    return images, torch.from_numpy(np.stack(projs)), dv, pairs


def run_scan(args, bench):
    import torch.distributed as dist
    from scene_3dreconstruction_mvsnet_b200 import _lib
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet
    from scene_3dreconstruction_mvsnet_b200.runner import DepthMapRunner, ScanRunner, plan_scan

    rank, world, local, dev = bench.init_dist()
    _lib.load()
    images, projs, dv, pairs = make_scan(seed=rank)
    images = images.pin_memory()
    torch.manual_seed(1)
    model = MVSNet(refine=False, precision="bf16").to(dev).eval()
    runner = ScanRunner(model, device=str(dev), pool_images=64)
    h, w = H_SCAN // 4, W_SCAN // 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: images resident in HBM; one step = FeatureNet over the 49 images + 49 depth maps
    d_images = images.to(dev)
    steps_plan = plan_scan(pairs, runner.pool_images)
    d_proj = [projs[s["views"]].unsqueeze(0).to(dev) for s in steps_plan]
    d_dv = dv.unsqueeze(0).to(dev)
    pool = torch.empty((64, h, 4, w, 8), dtype=torch.float16, device=dev)

    def device_scan():
        out = None
        for s, p in zip(steps_plan, d_proj):
            for img, slot in s["load"]:
                model.features_to_pool(d_images[img:img + 1], pool[slot:slot + 1])
            out = model.forward_from_pool(pool, s["slots"], p, d_dv)
        return out

    K, Wm = max(1, args.steps), max(1, min(args.warmup, 3))
    with torch.no_grad():
        for _ in range(Wm):
            device_scan()
        barrier()
        sampler = bench.ClockSampler(local)
        sampler.start()
        model.stage_events = []
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            device_scan()
        e1.record()
        barrier()
        launches = _lib.launch_count() - n0
        clocks = sampler.result()
        stage_ms = {}
        for marks in model.stage_events:
            for (n_a, a), (n_b, b) in zip(marks[:-1], marks[1:]):
                stage_ms.setdefault(n_b, []).append(a.elapsed_time(b))
        model.stage_events = None
        stage_ms = {k: sum(v) / len(v) for k, v in stage_ms.items()}
    ms = allmax(e0.elapsed_time(e1))
    value = world * K * NV / (ms * 1e-3)
    del d_images
    torch.cuda.empty_cache()

    # ---- e2e: ScanRunner.run_scan from pinned float32 host images (every image uploaded once per scan)
    acc = [0.0]

    def sink(k, d, c):
        acc[0] += float(d[0, 0, 0]) + float(c[0, 0, 0])

    def e2e(host_images):
        runner.run_scan(host_images, projs, dv, pairs, sink)  # warm-up (allocations)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            runner.run_scan(host_images, projs, dv, pairs, sink)
        barrier()
        t = allmax(time.perf_counter() - t0)
        return world * K * NV / t, runner.h2d_bytes, runner.d2h_bytes

    e2e_v, h2d, d2h = e2e(images)
    u8 = (images * 255.0).round().to(torch.uint8).pin_memory()
    e2e_u8, h2d_u8, _ = e2e(u8)
    del u8

    # ---- the per-call API on the same reference views (every view re-uploaded and re-extracted, like the reference)
    per_call = None
    if world == 1 and not args.quick:
        n = 10
        pc = DepthMapRunner(model, device=str(dev))
        views = [(images[[r] + s].unsqueeze(0).pin_memory(), projs[[r] + s].unsqueeze(0).pin_memory(),
                  dv.unsqueeze(0).pin_memory()) for r, s in pairs[:n]]
        pc.run_views(views[:3], sink)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        pc.run_views(views, sink)
        torch.cuda.synchronize(dev)
        per_call = {"value": n / (time.perf_counter() - t0), "unit": bench.UNIT, "reference_views": n,
                    "h2d_bytes_per_depth_map": pc.h2d_bytes_per_view,
                    "api": "DepthMapRunner.run_views: 5 images uploaded and FeatureNet on 5 images per depth map"}
        del pc, views

    if rank == 0:
        hbm_peak, tf_peak, peak_kind = bench.measured_peaks()
        wv = stage_ms.get("warp_variance")
        alg = 2 * 32 * D_SCAN * h * w + 4 * V_SCAN * 32 * h * w
        line = {
            "metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "step": "one scan = %d reference views (depth maps) over %d images" % (NV, NV),
            "ms_per_depth_map": ms / K / NV, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": bench.TC_DTYPE, "data": "synthetic", "config": bench.base_config("c5_scan"), "clocks": clocks,
            "e2e": {"value": e2e_v, "unit": bench.UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "h2d_bytes_per_depth_map": h2d / NV,
                    "api": "ScanRunner.run_scan (pinned float32 host images; each image uploaded and passed through "
                           "FeatureNet once per scan, features pooled on the device; depth+confidence D2H per view)"},
            "e2e_uint8_images": {"value": e2e_u8, "unit": bench.UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h},
            "per_call_api": per_call,
            "gpu_launches": int(launches), "stage_ms": stage_ms,
            "roofline": {"kernel": "warp_variance_win_kernel (indexed feature pool)", "bound": "hbm", "ms": wv,
                         "algorithmic_bytes": alg, "achieved": alg / (wv * 1e-3) / 1e9 if wv else None, "peak": hbm_peak,
                         "unit": "GB/s", "frac": alg / (wv * 1e-3) / 1e9 / hbm_peak if wv else None, "traffic": None,
                         "peak_kind": peak_kind},
        }
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            cb = bench.reference_cpu_forward("c5_scan", steps=1, warmup=1, budget_s=60.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
# c4_train: forward + backward + Adam at 3 views 512x640, D = 192, batch 4 per GPU (reference: train.py:241-300)
# ------------------------------------------------------------------------------------------------------------------
def run_train(args, bench):
    import torch.distributed as dist
    from scene_3dreconstruction_mvsnet_b200 import _lib, ops, synth
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet, mvsnet_loss

    rank, world, local, dev = bench.init_dist()
    torch.backends.cudnn.benchmark = True
    _lib.load()
    B, V, H, W, D = 4, 3, 512, 640, 192
    h, w = H // 4, W // 4
    torch.manual_seed(1)
    model = MVSNet(refine=False).to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    imgs, proj, dv = synth.make_inputs(B=B, V=V, H=H, W=W, D=D, focal=361.5, interval_scale=1.06, seed=rank)
    gt = torch.full((B, h, w), 650.0)
    mask = torch.ones(B, h, w)
    host = [t.pin_memory() for t in (imgs, proj, dv, gt, mask)]
    d_in = [t.to(dev) for t in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step(inp):
        opt.zero_grad(set_to_none=True)
        out = net(inp[0], inp[1], inp[2])
        loss = mvsnet_loss(out["depth"], inp[3], inp[4])
        loss.backward()
        opt.step()
        return loss

    K, Wm = args.steps, max(args.warmup, 3)
    for _ in range(Wm):
        step(d_in)
    barrier()
    sampler = bench.ClockSampler(local)
    sampler.start()
    ops.KERNEL_EVENTS = []
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = step(d_in)
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    clocks = sampler.result()
    kern_ms = {}
    for name, a, b in ops.KERNEL_EVENTS:
        kern_ms.setdefault(name, []).append(a.elapsed_time(b))
    ops.KERNEL_EVENTS = None
    kern_ms = {k: sum(v) / len(v) for k, v in kern_ms.items()}
    ms = allmax(e0.elapsed_time(e1))
    value = world * B * K / (ms * 1e-3)

    # e2e: every step uploads its batch from pinned host memory and reads the loss back
    def e2e_step():
        inp = [t.to(dev, non_blocking=True) for t in host]
        return float(step(inp).item())

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        last = e2e_step()
    barrier()
    e2e_v = world * B * K / allmax(time.perf_counter() - t0)
    h2d = sum(t.numel() * t.element_size() for t in host)

    # DDP's gradient all-reduce on its own (one 1.35 MB bucket): the collective this workload adds
    allreduce_ms = None
    if world > 1:
        g = torch.zeros(sum(p.numel() for p in model.parameters()), device=dev)
        for _ in range(5):
            dist.all_reduce(g)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dist.all_reduce(g)
        b.record()
        torch.cuda.synchronize(dev)
        allreduce_ms = a.elapsed_time(b) / 20

    eager = None
    if world == 1 and not args.quick:
        # the reference's stock eager-CUDA training step on this GPU (same ATen/cuDNN calls: oracle/torch_port.py)
        from oracle import torch_port
        del net, opt
        torch.cuda.empty_cache()
        sd = {k: v.to(dev) for k, v in bench.seeded_state_dict().items()}
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
        ropt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-3)

        def rstep():
            ropt.zero_grad(set_to_none=True)
            out = torch_port.mvsnet_forward_train(d_in[0], d_in[1], d_in[2], sd)
            m = d_in[4] > 0.5
            l = torch.nn.functional.smooth_l1_loss(out["depth"][m], d_in[3][m])
            l.backward()
            ropt.step()

        torch.cuda.reset_peak_memory_stats(dev)
        for _ in range(2):
            rstep()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            rstep()
        b.record()
        torch.cuda.synchronize(dev)
        rms = a.elapsed_time(b) / 3
        eager = {"value": B / (rms * 1e-3), "unit": "samples/s", "ms_per_step": rms, "steps": 3, "warmup": 2,
                 "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9,
                 "what": "oracle/torch_port.mvsnet_forward_train (the reference's ATen/cuDNN calls, train mode) + "
                         "smooth-L1 + backward + Adam on CUDA tensors, PyTorch defaults"}

    if rank == 0:
        hbm_peak, _, peak_kind = bench.measured_peaks()
        bwd_bytes = 4 * B * 32 * D * h * w + 3 * 4 * B * V * 32 * h * w     # SURVEY 8(d) kernel 4
        fwd_bytes = 4 * B * 32 * D * h * w + 4 * B * V * 32 * h * w
        bwd_ms, fwd_ms = kern_ms.get("warp_variance_bwd"), kern_ms.get("warp_variance_fwd")
        line = {
            "metric": "training samples/s (3 views 512x640, D=192, fwd+bwd+Adam)", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (TF32 allowed in cuDNN: PyTorch's default, as in the reference)",
            "data": "synthetic", "config": bench.base_config("c4_train"), "clocks": clocks, "loss": float(loss.detach()),
            "e2e": {"value": e2e_v, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "api": "MVSNet(...).train() forward + mvsnet_loss + backward + Adam, batch uploaded from pinned host "
                           "memory and the loss read back every step"},
            "gpu_launches": int(launches),
            "our_kernel_ms_per_step": kern_ms,
            "roofline": {"kernel": "warp_volume_bwd_kernel (backward of the fused warp+variance, kernel 4)", "bound": "hbm",
                         "ms": bwd_ms, "algorithmic_bytes": bwd_bytes,
                         "achieved": bwd_bytes / (bwd_ms * 1e-3) / 1e9 if bwd_ms else None, "peak": hbm_peak, "unit": "GB/s",
                         "frac": bwd_bytes / (bwd_ms * 1e-3) / 1e9 / hbm_peak if bwd_ms else None, "traffic": None,
                         "peak_kind": peak_kind},
            "roofline_fwd": {"kernel": "warp_variance_win32_kernel (fp32)", "ms": fwd_ms, "algorithmic_bytes": fwd_bytes,
                             "frac": fwd_bytes / (fwd_ms * 1e-3) / 1e9 / hbm_peak if fwd_ms else None},
            "collective": {"what": "DDP gradient all-reduce (338,129 fp32 = 1.35 MB, one bucket) over NCCL",
                           "allreduce_ms": allreduce_ms},
            "cuda_eager_baseline": eager,
            "max_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
