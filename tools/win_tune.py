"""Times the TMA-window warp+variance kernel (fp16 texels -> bf16 CP8) at the DTU shape for its tuning knobs
(MVS_WIN_CONFIG, MVS_WIN_HACC, MVS_WARP_DCHUNK).
Usage: python tools/win_tune.py [yaw]"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from scene_3dreconstruction_mvsnet_b200 import ops, synth
    yaw = float(sys.argv[2])
    V, H, W, D = 5, 1152, 1600, 192
    fea = synth.make_features(1, V, 32, H // 4, W // 4, seed=0).cuda()
    fea16 = fea.half().permute(0, 1, 3, 4, 2).contiguous()
    _, proj, dv = synth.make_inputs(B=1, V=V, H=H, W=W, D=D, focal=723.0, interval_scale=1.06, yaw=yaw)
    proj, dv = proj.cuda(), dv.cuda()
    def t(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    a = t(lambda: ops.warp_variance_cp8(fea16, proj, dv))
    b = t(lambda: ops.warp_variance_cp8(fea, proj, dv))
    print("cfg=%s hacc=%s dchunk=%s yaw=%g: fp16-nhwc in %.3f ms, fp32-nchw in %.3f ms (layout pass included)" % (
        os.environ.get("MVS_WIN_CONFIG", "0"), os.environ.get("MVS_WIN_HACC", "default"),
        os.environ.get("MVS_WARP_DCHUNK", "16"), yaw, a, b),
        flush=True)
else:
    yaw = sys.argv[1] if len(sys.argv) > 1 else "0"
    runs = [dict(MVS_WIN_HACC="0"), dict(MVS_WIN_HACC="1"), dict(MVS_WIN_CONFIG="1")]
    for dc in ("8", "16", "32"):
        runs.append(dict(MVS_WARP_DCHUNK=dc))
    for r in runs:
        subprocess.run([sys.executable, __file__, "child", yaw], env=dict(os.environ, **r))
