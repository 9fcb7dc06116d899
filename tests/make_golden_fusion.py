"""Generates tests/golden/fusion_*.npz from the UNMODIFIED reference (run in the build container, where
/root/reference and cv2 exist):  python tests/make_golden_fusion.py
The reference's eval.py cannot be imported as a module (argparse at import time, tensorboardX / open3d / plyfile are
not installed), so the three pure functions are taken from its AST and executed with numpy / cv2 -- nothing is copied
into this repository.  Also checks oracle.fusion_oracle.remap_bilinear bit-exactly against cv2.remap."""
import ast
import os
import sys
import types

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fusion_oracle as fo  # noqa: E402

REF = "/root/reference/eval.py"


def reference_functions(condmask_pixel, condmask_depth):
    src = open(REF).read()
    want = {"reproject_with_depth", "check_geometric_consistency"}
    code = [ast.get_source_segment(src, n) for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = {"np": np, "cv2": cv2, "args": types.SimpleNamespace(condmask_pixel=condmask_pixel, condmask_depth=condmask_depth),
          "print": lambda *a, **k: None}
    exec("\n\n".join(code), ns)
    return ns


def scene(h, w, nsrc, seed):
    """A tilted plane n.X = c (reference-camera coordinates) seen by nsrc+1 cameras: every depth map is the exact ray /
    plane intersection plus noise, outliers (reference) and holes (sources), so most pixels are geometrically consistent."""
    rs = np.random.RandomState(seed)
    f = 0.9 * w
    K = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1]], np.float64)
    n, c = np.array([0.12, -0.08, 1.0]), 600.0
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    pix = np.stack([xx, yy, np.ones_like(xx)], 0).reshape(3, -1).astype(np.float64)

    def plane_depth(Kc, E):
        R, t = E[:3, :3], E[:3, 3]
        rays = np.linalg.inv(Kc) @ pix                       # camera-frame rays with z = 1
        lam = (c + n @ (R.T @ t)) / (n @ (R.T @ rays))       # X_ref = R^T (lam * ray - t),  n.X_ref = c
        return lam.reshape(h, w).astype(np.float32)

    ref_depth = plane_depth(K, np.eye(4)) + (rs.randn(h, w) * 0.3).astype(np.float32)
    ref_depth[rs.rand(h, w) < 0.03] += 80.0  # outliers the filter must reject
    Ks, Es, Ds = [], [], []
    for s in range(nsrc):
        a = 0.05 * (s - nsrc / 2.0)
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        E = np.eye(4)
        E[:3, :3] = R
        E[:3, 3] = [-40.0 * (s + 1) * (1 if s % 2 else -1), 6.0 * s, 3.0 * s]
        Ksrc = K.copy()
        Ksrc[0, 0] *= 1.0 + 0.02 * s
        Ksrc[1, 1] *= 1.0 + 0.02 * s
        d = plane_depth(Ksrc, E) + (rs.randn(h, w) * 0.5).astype(np.float32)
        d[rs.rand(h, w) < 0.02] = 0.0  # holes
        Ks.append(Ksrc); Es.append(E); Ds.append(d)
    conf = rs.rand(h, w).astype(np.float32)
    return ref_depth, conf, K, np.eye(4), Ds, Ks, Es


def main():
    # 1. remap restatement vs cv2
    rs = np.random.RandomState(1)
    src = (rs.rand(37, 53) * 1000).astype(np.float32)
    mx = (rs.rand(64, 80) * 70 - 8).astype(np.float32)
    my = (rs.rand(64, 80) * 50 - 6).astype(np.float32)
    mx[0, 0], my[0, 1], mx[0, 2] = np.nan, np.inf, -1e20
    assert np.array_equal(cv2.remap(src, mx, my, interpolation=cv2.INTER_LINEAR), fo.remap_bilinear(src, mx, my))
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, (h, w, nsrc, seed, pix, dep, geo, photo) in {"fusion_a": (48, 64, 4, 0, 1.0, 0.01, 3, 0.8),
                                                           "fusion_b": (37, 53, 3, 1, 0.5, 0.005, 2, 0.3)}.items():
        ref_depth, conf, K, E, Ds, Ks, Es = scene(h, w, nsrc, seed)
        ns = reference_functions(pix, dep)
        masks, reps, xs, ys = [], [], [], []
        geo_sum = 0
        for d, k, e in zip(Ds, Ks, Es):
            m, dr, x2, y2 = ns["check_geometric_consistency"](ref_depth, K, E, d, k, e)
            masks.append(m); reps.append(dr); xs.append(x2); ys.append(y2)
            geo_sum = geo_sum + m.astype(np.int32)
        depth_avg = (sum(reps) + ref_depth) / (geo_sum + 1)            # eval.py:700
        geo_mask = geo_sum >= geo                                       # eval.py:702
        final = np.logical_and(conf > photo, geo_mask)                  # eval.py:660,703
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), ref_depth=ref_depth, conf=conf, K=K, E=E,
                            src_depths=np.stack(Ds), src_K=np.stack(Ks), src_E=np.stack(Es), masks=np.stack(masks),
                            reprojected=np.stack(reps), x_src=np.stack(xs), y_src=np.stack(ys), depth_avg=depth_avg,
                            geo_mask=geo_mask, final_mask=final, params=np.array([pix, dep, geo, photo], np.float64))
        print(name, "mask fractions", [float(m.mean()) for m in masks], "final", float(final.mean()))


if __name__ == "__main__":
    main()
