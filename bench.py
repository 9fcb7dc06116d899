#!/usr/bin/env python
"""bench.py -- depth maps/s of the MVSNet depth-inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json configs):
  c2_dtu_5view_1152x1600 (default)  one "step" = one depth map = one MVSNet.forward for one reference view (B=1) at
                                    the DTU eval shape: 5 views 1152x1600, D=192
  c1_3view_512x640, c3_bin_4view_512x640, *_rot       the same step at configs[0] / configs[2], rotated-camera variants
  c5_scan                           one DTU scan: 49 reference views x 5 views, every image uploaded and pushed through
                                    FeatureNet ONCE per scan (ScanRunner), sharded over the ranks by reference view
  c4_train                          configs[3]: forward + backward + Adam, 3 views 512x640, D=192, batch 4 per GPU,
                                    DDP over NCCL when N > 1; unit = samples/s
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.

  value      whole job over all ranks, inputs resident in HBM, CUDA-event timed, max over ranks, ONE stream (the same
             forwards over two streams: value_two_streams).  Every step's working
             set (1.4 GB cost volume) is far larger than the 126 MB L2, so no explicit L2 flush is needed (config.l2).
  e2e        same metric through the host-buffer API (DepthMapRunner.run_views): pinned host inputs, H2D copies and
             D2H reads of the results inside the timed region, every step.
  roofline   the fused warp+variance kernel (the "cost-vol HBM GB/s" half of BASELINE.json's metric): algorithmic
             bytes / CUDA-event time of the kernel inside the timed steps.
  cpu_baseline         the reference's CPU path on the host cores at the SAME configuration (full size, not a crop):
                       oracle/torch_port.py issues the same ATen CPU calls as the reference's MVSNet.forward.
  cuda_eager_baseline  the reference's stock eager-CUDA path on the same GPU: the same port with CUDA tensors (cuDNN +
                       ATen kernels), fp32 with TF32 off and with PyTorch's defaults.
  --impl reference     times the CPU path only, same config/metric/unit (/root/reference does not exist on the GPU box
                       and the reference is an unpackaged script repo: nothing to pip-install).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "depth maps/s (DTU 1152x1600, D=192, 5 views) @1/2/4/8 B200; cost-vol HBM GB/s"
UNIT = "depth maps/s"
DEFAULT_WORKLOAD = "c2_dtu_5view_1152x1600"
REFERENCE_BUDGET_S = 150.0   # --impl reference: wall-clock budget for warm-up + timed steps of the CPU path
TC_DTYPE = ("f16 operands / f32 accumulate in CostRegNet and FeatureNet (tcgen05, kind::f16); fp16 texels, packed-half tap "
            "interpolation and deviation sums in the fused warp kernel; fp16 cost volume and activations; f32 logits / softmax")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 6650.0, 1590.0, "fallback"  # /opt/skills/guides/B200_PROFILING.md


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
        }
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def base_config(workload):
    """The `config` object: identical keys and values in both arms (--impl ours / reference)."""
    from scene_3dreconstruction_mvsnet_b200 import synth
    if workload == "c5_scan":
        V, H, W, D = 5, 1152, 1600, 192
        extra = {"reference_views_per_scan": 49, "images_per_scan": 49}
    elif workload == "c4_train":
        V, H, W, D = 3, 512, 640, 192
        extra = {"batch_per_gpu": 4, "optimizer": "Adam"}
    else:
        V, H, W, D, _, _ = synth.config_of(workload)
        extra = {}
    cfg = {"workload": workload, "views": V, "image": [H, W], "depth_planes": D, "batch": 4 if workload == "c4_train" else 1,
           "feature_map": [H // 4, W // 4], "weights": "random-init (seed 1)", "inputs": "synthetic, seed = rank",
           "l2": "per-step working set (>= 1.4 GB cost volume) >> 126 MB L2; no flush needed",
           "sharding": "one reference view stream per rank, no data-path collective"}
    cfg.update(extra)
    return cfg


def seeded_state_dict():
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet
    torch.manual_seed(1)
    return {k: v.detach().clone() for k, v in MVSNet(refine=False).state_dict().items()}


def reference_cpu_forward(workload, steps, warmup, budget_s, seed=0):
    """The reference's CPU path (oracle/torch_port.py: the same ATen calls as /root/reference/models) at the workload's
    FULL size on all host threads.  Runs `warmup` untimed and up to `steps` timed forwards, stopping early when the
    wall-clock budget is used up; reports what it actually ran."""
    from oracle import torch_port
    from scene_3dreconstruction_mvsnet_b200 import synth
    torch.set_num_threads(os.cpu_count())
    sd = seeded_state_dict()
    name = DEFAULT_WORKLOAD if workload == "c5_scan" else workload
    imgs, proj, dv = synth.make_named(name, B=1, seed=seed)
    t_start = time.perf_counter()
    times = []
    warm_run = 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        torch_port.mvsnet_forward(imgs, proj, dv, sd)
        dt = time.perf_counter() - t0
        if i < warmup:
            warm_run += 1
        else:
            times.append(dt)
        spent = time.perf_counter() - t_start
        if times and spent + 1.2 * dt > budget_s:
            break
        if not times and i + 1 >= warmup and spent > 0.6 * budget_s:
            warmup = i + 1  # the warm-up alone used most of the budget: time at least one step
    t = sum(times) / len(times)
    V, H, W, D, _, _ = synth.config_of(name)
    return {"value": 1.0 / t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle/torch_port.py (the reference's own ATen CPU calls) at the full workload size %dx%d, V=%d, "
                      "D=%d, B=1: %d timed forward(s) after %d warm-up, %.2f s each" % (H, W, V, D, len(times), warm_run, t),
            "steps_run": len(times), "warmup_run": warm_run, "ms_per_step": 1e3 * t}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "c4_train":
        return run_reference_train(args)
    cb = reference_cpu_forward(args.workload, args.steps, min(args.warmup, 1), REFERENCE_BUDGET_S)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": cb["steps_run"], "warmup": cb["warmup_run"], "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args.workload),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "steps/warmup are what was actually run inside a %d s budget (one forward of this config takes several "
                "seconds on the host cores); requested values are in *_requested" % int(REFERENCE_BUDGET_S),
    }
    print(json.dumps(line), flush=True)


def run_reference_train(args):
    """configs[3] on the host cores: the reference's training step (train.py:241-300) = forward + smooth-L1 loss +
    backward + Adam through the reference's own ATen calls (autograd over oracle/torch_port.py's functional port)."""
    from oracle import torch_port
    from scene_3dreconstruction_mvsnet_b200 import synth
    torch.set_num_threads(os.cpu_count())
    B = 1  # B=4 needs ~40 GB of autograd state on the host; B=1 steps are timed and samples/s reported
    sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
          for k, v in seeded_state_dict().items()}
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)
    imgs, proj, dv = synth.make_inputs(B=B, V=3, H=512, W=640, D=192, focal=361.5, interval_scale=1.06, seed=0)
    gt = torch.full((B, 128, 160), 650.0)
    times = []
    t_start = time.perf_counter()
    for i in range(1 + args.steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = torch_port.mvsnet_forward_train(imgs, proj, dv, sd)
        loss = torch.nn.functional.smooth_l1_loss(out["depth"], gt)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= 1:
            times.append(dt)
        if times and time.perf_counter() - t_start + 1.2 * dt > REFERENCE_BUDGET_S:
            break
    t = sum(times) / len(times)
    line = {"impl": "reference", "metric": "training samples/s (3 views 512x640, D=192, fwd+bwd+Adam)", "value": B / t,
            "unit": "samples/s", "n_gpus": args.gpus, "steps": len(times), "warmup": 1, "ms_per_step": 1e3 * t,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config("c4_train"),
            "cpu_baseline": {"value": B / t, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "B=1 training steps (forward+backward+Adam) of the torch port on the host cores"},
            "e2e": {"value": B / t, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cuda_eager_forward(workload, dev, d_imgs, d_proj, d_dv, allow_tf32, steps=3, warmup=2):
    """The reference's stock eager-CUDA path on this GPU: oracle/torch_port.py with CUDA tensors (the same ATen /
    cuDNN calls the reference's MVSNet.forward makes after .cuda()), cudnn.benchmark on like eval.py:24."""
    from oracle import torch_port
    sd = {k: v.to(dev) for k, v in seeded_state_dict().items()}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = allow_tf32
    try:
        for _ in range(warmup):
            torch_port.mvsnet_forward(d_imgs, d_proj, d_dv, sd)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = torch_port.mvsnet_forward(d_imgs, d_proj, d_dv, sd)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    del sd
    torch.cuda.empty_cache()
    return {"value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "cudnn_allow_tf32": bool(allow_tf32), "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9,
            "depth_mean": float(out["depth"].mean())}


def init_dist():
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, local, dev


def run_ours(args):
    import torch.distributed as dist
    from scene_3dreconstruction_mvsnet_b200 import _lib, synth
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet
    from scene_3dreconstruction_mvsnet_b200.runner import DepthMapRunner

    rank, world, local, dev = init_dist()
    # strict fp32: every cuDNN call (none on our inference path; the eager baseline's) stays off TF32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True  # like the reference (eval.py:24)
    _lib.load()

    V, H, W, D, focal, itv = synth.config_of(args.workload)
    h, w = H // 4, W // 4
    # each rank owns its own shard of reference views (weak scaling): different seed per rank
    imgs, proj, dv = synth.make_named(args.workload, B=1, seed=rank)
    d_in = (imgs.to(dev), proj.to(dev), dv.to(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(precision, steps, warmup, inputs):
        """Device-resident throughput of MVSNet.forward: CUDA events around `steps` forwards, max over ranks."""
        torch.manual_seed(1)
        model = MVSNet(refine=False, precision=precision).to(dev).eval()
        with torch.no_grad():
            for _ in range(warmup):
                model(*inputs)
            barrier()
            sampler = ClockSampler(local)
            sampler.start()
            model.stage_events = []
            launches0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model(*inputs)
            e1.record()
            barrier()
            launches = _lib.launch_count() - launches0
            clocks = sampler.result()
            elapsed_ms = e0.elapsed_time(e1)
            stage_ms = {}
            for marks in model.stage_events:
                for (n0, a), (n1, b) in zip(marks[:-1], marks[1:]):
                    stage_ms.setdefault(n1, []).append(a.elapsed_time(b))
            model.stage_events = None
            stage_ms = {k: sum(v) / len(v) for k, v in stage_ms.items()}
        max_ms = allmax(elapsed_ms)
        return {"model": model, "value": world * steps / (max_ms * 1e-3), "max_ms": max_ms, "stage_ms": stage_ms,
                "launches": int(launches), "clocks": clocks}

    def roofline_of(precision, stage_ms):
        # volume written once + every feature map read once.  In the tensor-core mode the
        # fused kernel writes the volume as 16-bit (the tensor-core CostRegNet's input), i.e. 2 bytes per element.
        wv_ms = stage_ms.get("warp_variance")
        vol_elem = 4 if precision == "fp32" else 2
        alg_bytes = vol_elem * 32 * D * h * w + 4 * V * 32 * h * w
        achieved = alg_bytes / (wv_ms * 1e-3) / 1e9 if wv_ms else None
        return alg_bytes, wv_ms, achieved

    main = measure(args.precision, args.steps, args.warmup, d_in)
    model, value, max_ms, stage_ms, launches, clocks = (main["model"], main["value"], main["max_ms"], main["stage_ms"],
                                                        main["launches"], main["clocks"])

    # ---------------- same forwards alternating over two streams (independent depth maps overlap at kernel tails) -----
    value_2s = None
    if args.precision != "fp32":
        streams = [torch.cuda.Stream(dev) for _ in range(2)]
        with torch.no_grad():
            for st in streams:
                with torch.cuda.stream(st):
                    for _ in range(3):
                        model(*d_in)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for st in streams:
                st.wait_event(e0)
            for i in range(args.steps):
                with torch.cuda.stream(streams[i % 2]):
                    model(*d_in)
            for st in streams:
                torch.cuda.current_stream(dev).wait_stream(st)
            e1.record()
            barrier()
        value_2s = world * args.steps / (allmax(e0.elapsed_time(e1)) * 1e-3)
        del streams

    # ---------------- end to end through the host-buffer API ----------------
    runner = DepthMapRunner(model, device=str(dev))
    p_imgs, p_proj, p_dv = imgs.pin_memory(), proj.pin_memory(), dv.pin_memory()
    sink_acc = [0.0]

    def sink(i, depth_np, conf_np):
        sink_acc[0] += float(depth_np[0, 0, 0]) + float(conf_np[0, 0, 0])  # the host really reads the result

    def e2e_run(host_imgs):
        runner.run_views([(host_imgs, p_proj, p_dv)] * max(2, args.warmup), sink)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        runner.run_views([(host_imgs, p_proj, p_dv)] * args.steps, sink)
        e1.record()
        barrier()
        ms = allmax(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
        return world * args.steps / (ms * 1e-3)

    e2e_value = e2e_run(p_imgs)
    e2e_h2d = runner.h2d_bytes_per_view
    # same call with the images as 8-bit host buffers (as decoded from disk): /255 runs on the device, the upload is
    # 4x smaller.  Reported next to `e2e` (which keeps the reference loader's float32 images), not instead of it.
    u8_imgs = (imgs * 255.0).round().to(torch.uint8).pin_memory()
    e2e_u8_value = e2e_run(u8_imgs)
    e2e_u8_h2d = runner.h2d_bytes_per_view

    # ---- the scan-level API on a DTU-shaped scan (49 images, 49 reference views x 5 views): float32 host images, each
    # uploaded and passed through FeatureNet once per scan (ScanRunner; workload c5_scan has the full line)
    e2e_scan = None
    if args.workload == DEFAULT_WORKLOAD and args.precision != "fp32":
        import bench_workloads
        from scene_3dreconstruction_mvsnet_b200.runner import ScanRunner
        s_images, s_projs, s_dv, s_pairs = bench_workloads.make_scan(seed=rank)
        s_images = s_images.pin_memory()
        srunner = ScanRunner(model, device=str(dev), pool_images=64)
        srunner.run_scan(s_images, s_projs, s_dv, s_pairs, sink)
        n_scans = max(1, args.steps // 10)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_scans):
            srunner.run_scan(s_images, s_projs, s_dv, s_pairs, sink)
        barrier()
        t_scan = allmax(time.perf_counter() - t0)
        e2e_scan = {"value": world * n_scans * len(s_pairs) / t_scan, "unit": UNIT, "scans": n_scans,
                    "h2d_bytes_per_depth_map": srunner.h2d_bytes / len(s_pairs),
                    "d2h_bytes_per_depth_map": srunner.d2h_bytes / len(s_pairs),
                    "api": "ScanRunner.run_scan: one DTU-shaped scan per rank (49 float32 host images 1152x1600, 49 reference "
                           "views x 5 views), every image uploaded and passed through FeatureNet once per scan"}
        del srunner, s_images
        torch.cuda.empty_cache()

    if rank == 0:
        hbm_peak, tf_peak, peak_kind = measured_peaks()
        alg_bytes, wv_ms, achieved = roofline_of(args.precision, stage_ms)
        # DRAM bytes per launch of the same kernel at the same shape, from the committed ncu capture (not measured live)
        traffic = {}
        if args.workload == DEFAULT_WORKLOAD:
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                    traffic = json.load(f).get("bf16" if args.precision != "fp32" else "fp32", {}) or {}
            except (OSError, ValueError):
                traffic = {}
        cr_ms = stage_ms.get("cost_regularization")
        flops = 20304.0 * D * h * w
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else TC_DTYPE,
            "data": "synthetic",
            "config": base_config(args.workload),
            "implementation": {"precision": args.precision,
                               "featurenet": "fp32 FMA on own CUDA-core kernels (ops.featurenet_fp32, TMA halo tiles)" if args.precision == "fp32"
                               else "tcgen05 implicit GEMM, fp16 operands / f32 accumulate (ops.featurenet_tc)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d,
                    "d2h_bytes_per_step": runner.d2h_bytes_per_view, "api": "DepthMapRunner.run_views (pinned host "
                    "inputs -> H2D -> MVSNet.forward -> D2H depth+confidence, double-buffered)"},
            "e2e_uint8_images": {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": e2e_u8_h2d,
                                 "d2h_bytes_per_step": runner.d2h_bytes_per_view,
                                 "note": "same API, images as uint8 host buffers; /255 on the device (reference: on the host)"},
            "e2e_scan_api": e2e_scan,
            "value_two_streams": {"value": value_2s, "unit": UNIT, "note": "the same K forwards alternating over two CUDA "
                                  "streams (what DepthMapRunner does): the next depth map's kernels fill the SMs a kernel's "
                                  "tail leaves idle; `value` and the stage times are single-stream"},
            "gpu_launches": int(launches),
            "stage_ms": stage_ms,
            "roofline": {"kernel": ("warp_variance_win32_kernel" if args.precision == "fp32" else "warp_variance_win_kernel") +
                                   " (+ homography compose, ~0.5% of the stage)",
                         "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak if achieved else None, "traffic": traffic.get("dram_bytes"),
                         "traffic_source": traffic.get("capture"), "traffic_measured_live": False, "peak_kind": peak_kind,
                         "algorithmic_bytes": alg_bytes, "ms": wv_ms},
            "roofline_costreg": {"kernel": "CostRegNet (11 fused conv launches)", "flop": flops, "ms": cr_ms,
                                 "achieved": flops / (cr_ms * 1e-3) / 1e12 if cr_ms else None, "unit": "TFLOP/s",
                                 "bound": "fp32-fma" if args.precision == "fp32" else "tensor",
                                 "peak_bf16_tensor": tf_peak},
        }
    del runner
    if world == 1 and not args.quick:
        # ---- the same step with rotated (non-rectified) cameras: the window planner's cost on DTU-like geometry
        rot = args.workload + "_rot"
        if rot in synth.ROTATED:
            r_in = tuple(t.to(dev) for t in synth.make_named(rot, B=1, seed=rank))
            del model
            torch.cuda.empty_cache()
            n = max(5, args.steps // 2)
            o = measure(args.precision, n, 3, r_in)
            rb, rms, rach = roofline_of(args.precision, o["stage_ms"])
            line["rotated_cameras"] = {"workload": rot, "value": o["value"], "unit": UNIT, "ms_per_step": o["max_ms"] / n,
                                       "stage_ms": o["stage_ms"], "steps": n,
                                       "roofline_frac": rach / hbm_peak if rach else None,
                                       "cameras": "source views yawed 0.1 rad per 60 mm of baseline towards the scene, "
                                                  "rolled -8 / 5 / -10 / 7 degrees"}
            model = o["model"]
            del o, r_in
        # ---- the strict-fp32 mode (the reference's own precision), a few steps
        other = "fp32" if args.precision != "fp32" else "bf16"
        del model
        torch.cuda.empty_cache()
        n = max(3, args.steps // 4)
        o = measure(other, n, 3, d_in)
        ob, oms, oach = roofline_of(other, o["stage_ms"])
        line["other_precision_modes"] = [{"precision": other, "value": o["value"], "unit": UNIT,
                                          "ms_per_step": o["max_ms"] / n, "stage_ms": o["stage_ms"],
                                          "gpu_launches": o["launches"],
                                          "roofline": {"kernel": "warp_variance_win32_kernel" if other == "fp32" else "warp_variance_win_kernel",
                                                       "bound": "hbm", "algorithmic_bytes": ob, "ms": oms, "achieved": oach,
                                                       "unit": "GB/s", "frac": oach / hbm_peak if oach else None}}]
        del o
        torch.cuda.empty_cache()
        # ---- the reference's stock eager-CUDA path on this GPU ("the bar to beat on the B200 box")
        line["cuda_eager_baseline"] = {
            "what": "oracle/torch_port.py with CUDA tensors = the ATen/cuDNN calls of the reference's MVSNet.forward "
                    "(eval mode, no_grad, cudnn.benchmark), same inputs and weights, CUDA-event timed",
            "fp32_tf32_off": cuda_eager_forward(args.workload, dev, *d_in, allow_tf32=False),
            "pytorch_defaults": cuda_eager_forward(args.workload, dev, *d_in, allow_tf32=True)}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            cb = reference_cpu_forward(args.workload, steps=1, warmup=1, budget_s=60.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16", "tc", "fast"],
                    help="tc (= bf16, fast; default): FeatureNet and CostRegNet on the tcgen05 tensor cores with 16-bit "
                         "operands and fp32 accumulation, fp16 texels and packed-half arithmetic in the fused warp kernel "
                         "(the looser, separately stated tolerance north_star allows); fp32: strict CUDA-core path")
    ap.add_argument("--quick", action="store_true", help="main measurement and e2e only (no rotated-camera, fp32-mode, "
                                                         "eager-CUDA or CPU legs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c4_train":
        import bench_workloads
        bench_workloads.run_train(args, sys.modules[__name__])
    elif args.workload == "c5_scan":
        import bench_workloads
        bench_workloads.run_scan(args, sys.modules[__name__])
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
