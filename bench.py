#!/usr/bin/env python
"""bench.py -- depth maps/s of the MVSNet depth-inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one depth map = one MVSNet.forward for one reference view (B=1) of the workload
(default: BASELINE.json configs[1], the DTU eval shape: 5 views 1152x1600, D=192).  Prints ONE JSON
line (rank 0).  See DESIGN.md section "Measurement" for every field.

  value      depth maps/s, whole job over all ranks, inputs resident in HBM, CUDA-event timed,
             max over ranks.  Every step's working set (2.8 GB cost volume) is far larger than the
             126 MB L2, so no explicit L2 flush is needed between steps (config.l2).
  e2e        same metric through the host-buffer API (DepthMapRunner.run_views): pinned host inputs,
             H2D copies and D2H reads of the results inside the timed region, every step.
  roofline   the fused warp+variance kernel (the "cost-vol HBM GB/s" half of BASELINE.json's metric):
             algorithmic bytes / CUDA-event time of the kernel inside the timed steps.
  cpu_baseline  the oracle's torch-CPU port (same ATen calls as the reference) on the host cores.
  --impl reference   times that CPU port only (the reference is pure Python/PyTorch; its CPU path
             is these calls; /root/reference does not exist on the GPU box).
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
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_port_depth_maps_per_s(workload, steps, warmup, seed=0):
    """Times the oracle's torch-CPU port on a bounded sample: a 1/16-area crop of the workload
    (same V and D), all host threads.  depth maps/s = 1 / (16 * t_crop)."""
    from oracle import torch_port
    from scene_3dreconstruction_mvsnet_b200 import synth
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet
    V, H, W, D, focal, itv = synth.CONFIGS[workload]
    area_div = 16
    Hc, Wc = H // 4 // 32 * 32, W // 4 // 32 * 32
    area_div = (H * W) / float(Hc * Wc)
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(1)
    sd = {k: v.detach().clone() for k, v in MVSNet(refine=False).state_dict().items()}
    imgs, proj, dv = synth.make_inputs(B=1, V=V, H=Hc, W=Wc, D=D, focal=focal / 4, interval_scale=itv, seed=seed)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        torch_port.mvsnet_forward(imgs, proj, dv, sd)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return {"value": 1.0 / (area_div * t), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle/torch_port.py (same ATen CPU calls as the reference) on a 1/%.1f-area crop %dx%d of the "
                      "workload, V=%d D=%d, %d timed steps after %d warm-up; depth maps/s = 1/(%.1f*t_step), "
                      "t_step=%.3fs" % (area_div, Hc, Wc, V, D, steps, warmup, area_div, t),
            "t_step_s": t, "ms_per_depth_map": 1e3 * area_div * t}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_port_depth_maps_per_s(args.workload, args.steps, args.warmup)
    V, H, W, D, _, _ = __import__("scene_3dreconstruction_mvsnet_b200.synth", fromlist=["CONFIGS"]).CONFIGS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_depth_map"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "views": V, "image": [H, W], "depth_planes": D, "batch": 1},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    from scene_3dreconstruction_mvsnet_b200 import _lib, synth
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet
    from scene_3dreconstruction_mvsnet_b200.runner import DepthMapRunner

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # strict fp32: FeatureNet (cuDNN) is kept off TF32 so the whole depth map is an fp32 result
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True  # like the reference (eval.py:24)
    _lib.load()

    V, H, W, D, focal, itv = synth.CONFIGS[args.workload]
    h, w = H // 4, W // 4
    # each rank owns its own shard of reference views (weak scaling): different seed per rank
    imgs, proj, dv = synth.make_named(args.workload, B=1, seed=rank)
    d_imgs, d_proj, d_dv = imgs.to(dev), proj.to(dev), dv.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def measure(precision, steps, warmup):
        """Device-resident throughput of MVSNet.forward: CUDA events around `steps` forwards, max over ranks."""
        torch.manual_seed(1)
        model = MVSNet(refine=False, precision=precision).to(dev).eval()
        with torch.no_grad():
            for _ in range(warmup):
                model(d_imgs, d_proj, d_dv)
            barrier()
            sampler = ClockSampler(local)
            sampler.start()
            model.stage_events = []
            launches0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model(d_imgs, d_proj, d_dv)
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
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        max_ms = float(t.item())
        return {"model": model, "value": world * steps / (max_ms * 1e-3), "max_ms": max_ms, "stage_ms": stage_ms,
                "launches": int(launches), "clocks": clocks}

    main = measure(args.precision, args.steps, args.warmup)
    model, value, max_ms, stage_ms, launches, clocks = (main["model"], main["value"], main["max_ms"], main["stage_ms"],
                                                        main["launches"], main["clocks"])

    # ---------------- end to end through the host-buffer API ----------------
    runner = DepthMapRunner(model, device=str(dev))
    p_imgs, p_proj, p_dv = imgs.pin_memory(), proj.pin_memory(), dv.pin_memory()
    sink_acc = [0.0]

    def sink(i, depth_np, conf_np):
        sink_acc[0] += float(depth_np[0, 0, 0]) + float(conf_np[0, 0, 0])  # the host really reads the result

    runner.run_views([(p_imgs, p_proj, p_dv)] * max(2, args.warmup), sink)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    runner.run_views([(p_imgs, p_proj, p_dv)] * args.steps, sink)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([max(e2e_ms, e2e_wall_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / (float(t.item()) * 1e-3)

    # same call with the images as 8-bit host buffers (as decoded from disk): /255 runs on the device, the upload is
    # 4x smaller.  Reported next to `e2e` (which keeps the reference loader's float32 images), not instead of it.
    u8_imgs = (imgs * 255.0).round().to(torch.uint8).pin_memory()
    runner.run_views([(u8_imgs, p_proj, p_dv)] * max(2, args.warmup), sink)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    runner.run_views([(u8_imgs, p_proj, p_dv)] * args.steps, sink)
    e1.record()
    barrier()
    t = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_u8_value = world * args.steps / (float(t.item()) * 1e-3)
    e2e_u8_h2d = runner.h2d_bytes_per_view
    runner.run_views([(p_imgs, p_proj, p_dv)], sink)  # back to the float32 slots (h2d_bytes_per_view below refers to them)

    if rank == 0:
        hbm_peak, tf_peak, peak_kind = measured_peaks()
        wv_ms = stage_ms.get("warp_variance")
        # volume written once + every feature map read once.  In the bf16 mode the fused
        # kernel writes the volume as bf16 (the tensor-core CostRegNet's input), i.e. 2 bytes per element.
        vol_elem = 4 if args.precision == "fp32" else 2
        alg_bytes = vol_elem * 32 * D * h * w + 4 * V * 32 * h * w
        achieved = alg_bytes / (wv_ms * 1e-3) / 1e9 if wv_ms else None
        # DRAM bytes per launch of the same kernel at the same shape, from the committed ncu capture (not measured live)
        traffic = {}
        if args.workload == "c2_dtu_5view_1152x1600":
            try:
                with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
                    traffic = json.load(f).get(args.precision, {}) or {}
            except (OSError, ValueError):
                traffic = {}
        cr_ms = stage_ms.get("cost_regularization")
        flops = 20304.0 * D * h * w
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": {"fp32": "f32",
                      "bf16": "bf16 operands / f32 accumulate in CostRegNet, fp16 operands / f32 accumulate in FeatureNet (both tcgen05); fp16 texels + packed-half tap interpolation and deviation sums in the fused warp kernel",
                      "fast": "as bf16, FeatureNet in fp16 as well"}[args.precision],
            "data": "synthetic",
            "config": {"workload": args.workload, "views": V, "image": [H, W], "depth_planes": D, "batch": 1,
                       "feature_map": [h, w], "weights": "random-init (seed 1), eval mode",
                       "precision": args.precision,
                       "featurenet": {"fp32": "cuDNN NHWC fused conv+bias+relu, fp32 (TF32 off)",
                                      "bf16": "tcgen05 implicit GEMM, fp16 operands / f32 accumulate (ops.featurenet_tc)",
                                      "fast": "tcgen05 implicit GEMM, fp16 operands / f32 accumulate (ops.featurenet_tc)"}[args.precision],
                       "l2": "per-step working set (1.4-2.8 GB cost volume) >> 126 MB L2; no flush needed",
                       "sharding": "one reference view stream per rank, no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": runner.h2d_bytes_per_view,
                    "d2h_bytes_per_step": runner.d2h_bytes_per_view, "api": "DepthMapRunner.run_views (pinned host "
                    "inputs -> H2D -> MVSNet.forward -> D2H depth+confidence, double-buffered)"},
            "e2e_uint8_images": {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": e2e_u8_h2d,
                                 "d2h_bytes_per_step": runner.d2h_bytes_per_view,
                                 "note": "same API, images as uint8 host buffers; /255 on the device (reference: on the host)"},
            "gpu_launches": int(launches),
            "stage_ms": stage_ms,
            "roofline": {"kernel": ("warp_variance_fwd2_kernel" if args.precision == "fp32" else "warp_variance_win_kernel") +
                                   " (+ homography compose and feature layout pre-passes, ~2% of the stage)",
                         "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak if achieved else None, "traffic": traffic.get("dram_bytes"),
                         "traffic_source": traffic.get("capture"), "peak_kind": peak_kind,
                         "algorithmic_bytes": alg_bytes, "ms": wv_ms},
            "roofline_costreg": {"kernel": "CostRegNet (11 fused conv launches)", "flop": flops, "ms": cr_ms,
                                 "achieved": flops / (cr_ms * 1e-3) / 1e12 if cr_ms else None, "unit": "TFLOP/s",
                                 "bound": "fp32-fma" if args.precision == "fp32" else "tensor",
                                 "peak_bf16_tensor": tf_peak},
        }
        if world == 1 and not args.no_other_mode:
            del model, runner
            line["other_precision_modes"] = []
            for other in ("fp32", "bf16", "fast"):
                if other == args.precision:
                    continue
                torch.cuda.empty_cache()
                n = max(3, args.steps // 4)
                o = measure(other, n, 3)
                line["other_precision_modes"].append({"precision": other, "value": o["value"], "unit": UNIT,
                                                      "ms_per_step": o["max_ms"] / n, "stage_ms": o["stage_ms"],
                                                      "gpu_launches": o["launches"]})
                del o
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_port_depth_maps_per_s(args.workload, steps=2, warmup=1)
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
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16", "fast"],
                    help="bf16 (default): CostRegNet on tcgen05 tensor cores, the bf16 conv3d path north_star allows, fp32 "
                         "arithmetic elsewhere; fp32: strict CUDA-core path; fast: bf16 + fp16 features")
    ap.add_argument("--no-other-mode", action="store_true", help="skip the short run of the other precision mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
