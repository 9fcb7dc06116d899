"""GPU parity of the geometric-consistency filter (csrc/fusion.cu) with the reference, through the goldens generated
from the unmodified reference functions (tests/make_golden_fusion.py) and, at the DTU depth-map size, with the numpy
oracle on the same inputs.

Tolerance.  The chain is float64 with float32 casts exactly where the reference casts; the only freedom is the
summation order inside numpy's matmul (BLAS) versus the kernel's left-to-right products, i.e. differences of a few
float64 ulps before a cast.  A cast, a 1/32-pixel rounding inside cv2.remap or a threshold can amplify that at a tie, so
the comparison allows a tiny fraction of disagreeing pixels (<= 2e-4) and otherwise demands: masks identical,
reprojected depth and sample positions within 2 float32 ulps, averaged depth within 1e-6 relative."""
import os

import numpy as np
import pytest

from oracle import fusion_oracle as fo
from scene_3dreconstruction_mvsnet_b200 import fusion

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def close_f32(a, b, ulps=2):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    tol = ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32))
    return (np.abs(a - b) <= tol) | (np.isnan(a) & np.isnan(b))


def compare(out, masks, reps, xs, ys, avg, geo, final, max_bad=2e-4):
    n = masks.size
    bad_mask = (out["mask"] != masks)
    assert bad_mask.sum() <= max_bad * n + 1, "masks differ at %d / %d pixels" % (bad_mask.sum(), n)
    ok = ~bad_mask
    assert (close_f32(out["depth_reprojected"], reps) | ~ok).mean() > 1 - max_bad
    assert close_f32(out["x2d_src"], xs).mean() > 1 - max_bad and close_f32(out["y2d_src"], ys).mean() > 1 - max_bad
    pix_ok = ok.all(0)
    rel = np.abs(out["depth_est_averaged"] - avg) / np.maximum(np.abs(avg), 1e-12)
    assert (rel[pix_ok] < 1e-6).mean() > 1 - max_bad
    assert (out["geo_mask"] != geo).sum() <= max_bad * geo.size + 1 and (out["final_mask"] != final).sum() <= max_bad * geo.size + 1


@pytest.mark.parametrize("name", ["fusion_a", "fusion_b"])
def test_filter_view_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    pix, dep, geo, photo = g["params"]
    out = fusion.filter_view(g["ref_depth"], g["conf"], g["K"], g["E"], g["src_depths"], g["src_K"], g["src_E"], photomask=photo,
                             geomask=int(geo), condmask_pixel=pix, condmask_depth=dep, details=True)
    assert out["depth_est_averaged"].dtype == np.float64 and out["geo_mask_sum"].dtype == np.int32
    compare(out, g["masks"], g["reprojected"], g["x_src"], g["y_src"], g["depth_avg"], g["geo_mask"], g["final_mask"])
    assert np.array_equal(out["photo_mask"], g["conf"] > np.float32(photo))
    # and the drop-in of eval.py:564 for a single pair
    m, dr, xs, ys = fusion.check_geometric_consistency(g["ref_depth"], g["K"], g["E"], g["src_depths"][0], g["src_K"][0],
                                                       g["src_E"][0], pix, dep)
    assert (m != g["masks"][0]).sum() <= 2 and close_f32(dr, g["reprojected"][0]).mean() > 0.999


def test_filter_view_dtu_size_against_oracle_and_edge_cases():
    """DTU depth-map size (288x400), 10 source views (NviewFilter default): kernel vs the numpy oracle, including zero /
    negative / NaN depths and a source camera whose points fall outside the image."""
    rs = np.random.RandomState(7)
    h, w, S = 288, 400, 10
    f = 0.9 * w
    K = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1]], np.float64)
    ref = (600 + 30 * rs.rand(h, w)).astype(np.float32)
    ref[0, :8] = [0.0, -5.0, np.nan, np.inf, 1e-30, 1e30, 425.0, 935.0]
    Ks, Es, Ds = [], [], []
    for s in range(S):
        a = 0.04 * (s - 4)
        E = np.eye(4)
        E[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        E[:3, 3] = [30.0 * (s - 5), 5.0 * s, 2.0 * s] if s < 9 else [5000.0, 0, 0]   # last view: everything out of frame
        d = (600 + 30 * rs.rand(h, w)).astype(np.float32)
        d[rs.rand(h, w) < 0.05] = 0
        Ks.append(K * np.array([[1 + 0.01 * s], [1 + 0.01 * s], [1]])); Es.append(E); Ds.append(d)
    conf = rs.rand(h, w).astype(np.float32)
    out = fusion.filter_view(ref, conf, K, np.eye(4), np.stack(Ds), np.stack(Ks), np.stack(Es), details=True)
    masks, reps, xs, ys = [], [], [], []
    for d, k, e in zip(Ds, Ks, Es):
        m, dr, x2, y2 = fo.check_geometric_consistency(ref, K, np.eye(4), d, k, e)
        masks.append(m); reps.append(dr); xs.append(x2); ys.append(y2)
    avg, pm, gm, fm, gs = fo.filter_view(ref, conf, K, np.eye(4), Ds, Ks, Es)
    with np.errstate(invalid="ignore"):
        compare(out, np.stack(masks), np.stack(reps), np.stack(xs), np.stack(ys), avg, gm, fm)
    assert not out["mask"][9].any()                       # out-of-frame view is never consistent
    assert not out["mask"][:, 0, :6].any()                # degenerate reference depths are never consistent
