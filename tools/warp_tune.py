"""Times the fused warp+variance kernel (bf16 CP8 and fp32 outputs) at the DTU shape for mapping variants."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from scene_3dreconstruction_mvsnet_b200 import ops, synth
    V, H, W, D = 5, 1152, 1600, 192
    fea = synth.make_features(1, V, 32, H // 4, W // 4, seed=0).cuda()
    _, proj, dv = synth.make_named("c2_dtu_5view_1152x1600")
    proj, dv = proj.cuda(), dv.cuda()
    def t(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    a = t(lambda: ops.warp_variance_cp8(fea, proj, dv))
    b = t(lambda: ops.warp_variance_fwd(fea, proj, dv))
    print("dchunk=%s  cp8 %.3f ms  fp32 %.3f ms" % (os.environ.get("MVS_WARP_DCHUNK"), a, b))
else:
    for dc in (4, 8, 16, 32, 64, 192):
        env = dict(os.environ, MVS_WARP_DCHUNK=str(dc))
        subprocess.run([sys.executable, __file__, "child"], env=env)
