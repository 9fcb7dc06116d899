"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/make_golden.py

The reference (/root/reference/models) is imported under the alias `ref_models` so that it does not
clash with this repo's own drop-in `models` package.  /root/reference does not
exist on the GPU box, so the vectors written here are committed and are what pins the oracle
(tests/test_oracle_golden.py) and, through it, the CUDA path (tests/test_gpu_parity.py).

Fixtures:
  weights_calibrated.npz  "BN-calibrated" state_dict: default seeded init, every BN
                          momentum=1.0, one train()-mode forward so running stats = batch stats, then eval().
                          With default init the logits have std 2e-5 and every output is constant,
                          which no test could tell apart; this fixture has logits std ~0.5.
  case_a.npz  B=1 V=3 64x96  D=16, small yaw  -> features, warped (view 1), variance, logits, prob, depth, conf
  case_b.npz  B=2 V=4 32x64  D=8,  per-sample depth ranges, grayscale imgs -> same stages
  case_bwd.npz  grads of the train-branch variance volume w.r.t. every view's features
  case_default_init.npz  default (uncalibrated) weights at case_a's inputs: depth == mean(depth_values)
  config_c1.npz / config_c3.npz / config_c2.npz   (python tests/make_golden.py --config-sized)
              the BASELINE.json shapes C1 (3 views 512x640), C3 (4 views gray 512x640) and C2 (5 views 1152x1600),
              D=192, inputs = synth.make_named(name, seed=0), calibrated weights: depth, conf, index_f of the
              unmodified reference's MVSNet.forward, plus the variance volume and the logits at 16 / 256 sampled
              pixels (all planes) so that the fused kernels are pinned to the reference at full size too.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True


def import_reference():
    pkg = types.ModuleType("ref_models")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["ref_models"] = pkg
    import ref_models.module  # noqa: F401
    import ref_models.mvsnet  # noqa: F401
    return sys.modules["ref_models.mvsnet"], sys.modules["ref_models.module"]


def calibrated_model(ref_mvsnet, imgs, proj, dv, seed=1):
    torch.manual_seed(seed)
    m = ref_mvsnet.MVSNet(refine=False)
    for mod in m.modules():
        if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            mod.momentum = 1.0
    m.train()
    with torch.no_grad():
        m(imgs, proj, dv)
    m.eval()
    return m


def run_stages(ref_mvsnet, ref_module, m, imgs, proj, dv):
    """Re-run the reference forward stage by stage with the reference's own functions."""
    out = {}
    with torch.no_grad():
        views = torch.unbind(imgs, 1)
        projs = torch.unbind(proj, 1)
        feats = [m.feature(v) for v in views]
        out["features"] = torch.stack(feats, 1)
        out["warped_v1"] = ref_module.homo_warping(feats[1], projs[1], projs[0], dv)
        D, V = dv.shape[1], len(views)
        s = feats[0].unsqueeze(2).repeat(1, 1, D, 1, 1)
        q = s ** 2
        for f, p in zip(feats[1:], projs[1:]):
            w = ref_module.homo_warping(f, p, projs[0], dv)
            s += w
            q += w.pow_(2)
        var = q.div_(V).sub_(s.div_(V).pow_(2))
        out["variance"] = var.clone()
        logits = m.cost_regularization(var).squeeze(1)
        out["logits"] = logits
        out["prob"] = torch.softmax(logits, 1)
        res = m(imgs, proj, dv)  # the real thing, end to end
        out["depth"] = res["depth"]
        out["conf"] = res["photometric_confidence"]
        out["index_f"] = ref_module.depth_regression(out["prob"], torch.arange(D, dtype=torch.float32))
    return {k: v.numpy() for k, v in out.items()}


def config_sized(names=("c1_3view_512x640", "c3_bin_4view_512x640", "c2_dtu_5view_1152x1600")):
    """Outputs of the unmodified reference at the BASELINE shapes (CPU; C2 takes ~1 min and ~15 GB)."""
    import hashlib
    from scene_3dreconstruction_mvsnet_b200 import synth

    ref_mvsnet, ref_module = import_reference()
    gold = os.path.join(HERE, "golden")
    with np.load(os.path.join(gold, "weights_calibrated.npz")) as z:
        sd = {k: torch.from_numpy(z[k]) for k in z.files}
    m = ref_mvsnet.MVSNet(refine=False)
    m.load_state_dict(sd)
    m.eval()
    torch.set_num_threads(os.cpu_count())
    for name in names:
        imgs, proj, dv = synth.make_named(name, B=1, seed=0)
        st = run_stages(ref_mvsnet, ref_module, m, imgs, proj, dv)
        h, w = st["depth"].shape[1:]
        g = np.random.default_rng(17)
        pix = np.sort(g.choice(h * w, 256, replace=False))
        py, px = pix // w, pix % w
        tag = name.split("_")[0]
        np.savez_compressed(
            os.path.join(gold, "config_%s.npz" % tag), name=np.array(name),
            imgs_sha1=np.array(hashlib.sha1(imgs.numpy().tobytes()).hexdigest()),
            depth=st["depth"], conf=st["conf"], index_f=st["index_f"],
            sample_yx=np.stack([py, px], 1).astype(np.int32),
            logits_samples=st["logits"][0][:, py, px].astype(np.float32),          # [D, 256]
            variance_samples=st["variance"][0][:, :, py[:16], px[:16]].astype(np.float32),  # [32, D, 16]
            features_samples=st["features"][0][:, :, py, px].astype(np.float32))   # [V, 32, 256]
        print("%s: depth [%.1f, %.1f] conf [%.3f, %.3f] logits std %.3g" % (
            tag, st["depth"].min(), st["depth"].max(), st["conf"].min(), st["conf"].max(), st["logits"].std()),
            flush=True)
        del st


def main():
    from scene_3dreconstruction_mvsnet_b200 import synth

    if "--config-sized" in sys.argv:
        return config_sized()

    ref_mvsnet, ref_module = import_reference()
    gold = os.path.join(HERE, "golden")
    os.makedirs(gold, exist_ok=True)

    # ---- case A + calibrated weights
    imgs, proj, dv = synth.make_inputs(B=1, V=3, H=64, W=96, D=16, focal=60.0, interval_scale=6.0, yaw=0.03, seed=3)
    m = calibrated_model(ref_mvsnet, imgs, proj, dv)
    sd = {k: v.numpy() for k, v in m.state_dict().items()}
    np.savez_compressed(os.path.join(gold, "weights_calibrated.npz"), **sd)
    st = run_stages(ref_mvsnet, ref_module, m, imgs, proj, dv)
    np.savez_compressed(os.path.join(gold, "case_a.npz"), imgs=imgs.numpy(), proj=proj.numpy(), dv=dv.numpy(), **st)
    print("case_a: logits std %.3g depth [%.1f, %.1f] conf [%.3f, %.3f] var max %.3g" % (
        st["logits"].std(), st["depth"].min(), st["depth"].max(), st["conf"].min(), st["conf"].max(),
        st["variance"].max()))

    # ---- default-init weights at the same inputs (the BASELINE "random-init" criterion)
    torch.manual_seed(1)
    m0 = ref_mvsnet.MVSNet(refine=False).eval()
    with torch.no_grad():
        r0 = m0(imgs, proj, dv)
    np.savez_compressed(os.path.join(gold, "case_default_init.npz"), depth=r0["depth"].numpy(),
                        conf=r0["photometric_confidence"].numpy(), seed=np.int64(1))

    # ---- case B: batch 2, 4 views, grayscale, per-sample depth ranges (same calibrated weights)
    imgs, proj, dv = synth.make_inputs(B=2, V=4, H=32, W=64, D=8, focal=40.0, interval_scale=12.0, yaw=0.0, seed=5,
                                       gray=True)
    st = run_stages(ref_mvsnet, ref_module, m, imgs, proj, dv)
    np.savez_compressed(os.path.join(gold, "case_b.npz"), imgs=imgs.numpy(), proj=proj.numpy(), dv=dv.numpy(), **st)
    print("case_b: logits std %.3g depth [%.1f, %.1f]" % (st["logits"].std(), st["depth"].min(), st["depth"].max()))

    # ---- backward of the train-branch variance volume (mvsnet.py:167-169,177) w.r.t. features
    B, V, C, h, w, D = 2, 3, 32, 8, 16, 8
    g = torch.Generator().manual_seed(11)
    fea = torch.randn(B, V, C, h, w, generator=g, requires_grad=True)
    _, proj, dv = synth.make_inputs(B=B, V=V, H=4 * h, W=4 * w, D=D, focal=30.0, interval_scale=16.0, yaw=0.05,
                                    seed=7)
    gvar = torch.randn(B, C, D, h, w, generator=g)
    feats = torch.unbind(fea, 1)
    projs = torch.unbind(proj, 1)
    s = feats[0].unsqueeze(2).repeat(1, 1, D, 1, 1)
    q = s ** 2
    for f, p in zip(feats[1:], projs[1:]):
        wv = ref_module.homo_warping(f, p, projs[0], dv)
        s = s + wv
        q = q + wv ** 2
    var = q.div_(V).sub_(s.div_(V).pow_(2))
    var.backward(gvar)
    np.savez_compressed(os.path.join(gold, "case_bwd.npz"), fea=fea.detach().numpy(), proj=proj.numpy(),
                        dv=dv.numpy(), grad_var=gvar.numpy(), variance=var.detach().numpy(),
                        grad_fea=fea.grad.numpy())
    for f in sorted(os.listdir(gold)):
        print(f, os.path.getsize(os.path.join(gold, f)))


if __name__ == "__main__":
    main()
