"""Pin the oracle (oracle/mvsnet_oracle.c and oracle/torch_port.py) against golden vectors produced
by the unmodified reference (tests/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle import torch_port as port


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


@pytest.mark.parametrize("case", ["case_a", "case_b"])
def test_c_oracle_homo_warping(case, request):
    c = request.getfixturevalue(case)
    out = orc.homo_warping(c["features"][:, 1], c["proj"][:, 1], c["proj"][:, 0], c["dv"])
    # fp32 rounding of the coordinate chain only (reference inverts in fp32, oracle in fp64)
    assert maxabs(out, c["warped_v1"]) < 2e-4
    assert np.mean(c["warped_v1"] == 0) > 0.01  # the fixture does exercise zero padding
    assert np.mean(c["warped_v1"] != 0) > 0.3


@pytest.mark.parametrize("case", ["case_a", "case_b"])
def test_c_oracle_variance(case, request):
    c = request.getfixturevalue(case)
    var = orc.warp_variance(c["features"], c["proj"], c["dv"])
    assert maxabs(var, c["variance"]) < 2e-4


@pytest.mark.parametrize("case", ["case_a", "case_b"])
def test_c_oracle_costreg(case, request, weights):
    c = request.getfixturevalue(case)
    logits = orc.cost_regularization(c["variance"], weights)[:, 0]
    assert c["logits"].std() > 0.1  # discriminating fixture
    assert maxabs(logits, c["logits"]) < 2e-4


@pytest.mark.parametrize("case", ["case_a", "case_b"])
def test_c_oracle_tail(case, request):
    c = request.getfixturevalue(case)
    depth, conf, idx, prob = orc.softmax_depth_conf(c["logits"], c["dv"], want_prob=True)
    rng = float(c["dv"].max() - c["dv"].min())
    assert maxabs(prob, c["prob"]) < 1e-6
    assert maxabs(depth, c["depth"]) < 1e-3 * rng
    assert maxabs(idx, c["index_f"]) < 1e-4
    # the 4-plane window moves when trunc(index) flips: compare away from integer boundaries
    safe = np.abs(c["index_f"] - np.round(c["index_f"])) > 1e-3
    assert safe.mean() > 0.9
    assert np.max(np.abs(conf - c["conf"])[safe]) < 1e-6


def test_c_oracle_depth_regression_1d(case_a):
    D = case_a["prob"].shape[1]
    out = orc.depth_regression(case_a["prob"], np.arange(D, dtype=np.float32))
    assert maxabs(out, case_a["index_f"]) < 1e-5
    out2 = orc.depth_regression(case_a["prob"], case_a["dv"])
    assert maxabs(out2, case_a["depth"]) < 1e-3


def test_c_oracle_backward(case_bwd):
    c = case_bwd
    var = orc.warp_variance(c["fea"], c["proj"], c["dv"])
    assert maxabs(var, c["variance"]) < 2e-4
    g = orc.warp_variance_bwd(c["grad_var"], c["fea"], c["proj"], c["dv"])
    assert g.shape == c["grad_fea"].shape
    scale = float(np.abs(c["grad_fea"]).max())
    assert maxabs(g, c["grad_fea"]) < 2e-4 * max(scale, 1.0)


@pytest.mark.parametrize("case", ["case_a", "case_b"])
def test_torch_port_end_to_end(case, request, weights):
    c = request.getfixturevalue(case)
    sd = {k: torch.from_numpy(v) for k, v in weights.items()}
    st = {}
    out = port.mvsnet_forward(torch.from_numpy(c["imgs"]), torch.from_numpy(c["proj"]), torch.from_numpy(c["dv"]),
                              sd, stages=st)
    assert maxabs(st["features"], c["features"]) < 1e-5
    assert maxabs(st["variance"], c["variance"]) < 1e-5
    assert maxabs(st["logits"], c["logits"]) < 1e-4
    assert maxabs(out["depth"], c["depth"]) < 1e-3
    assert maxabs(out["photometric_confidence"], c["conf"]) < 1e-4


def test_default_init_is_degenerate(case_a):
    """default init => uniform softmax, depth == mean(depth_values), conf == 4/D."""
    from conftest import load_golden
    d = load_golden("case_default_init.npz")
    dv = case_a["dv"]
    assert maxabs(d["depth"], np.full_like(d["depth"], dv.mean())) < 0.05
    assert maxabs(d["conf"], np.full_like(d["conf"], 4.0 / dv.shape[1])) < 1e-3


# ------------------------------------------------------------------------------------------------
# Geometric-consistency filter (the step after the depth path, reference eval.py:508-585, 660-703)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["fusion_a", "fusion_b"])
def test_fusion_oracle_matches_reference_golden(name):
    """oracle/fusion_oracle.py (numpy restatement incl. cv2.remap) against outputs of the unmodified reference functions
    (tests/make_golden_fusion.py).  Same float64 numpy operations: everything is bit-identical."""
    import os
    from oracle import fusion_oracle as fo
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    pix, dep, geo, photo = g["params"]
    for s in range(g["src_depths"].shape[0]):
        m, dr, xs, ys = fo.check_geometric_consistency(g["ref_depth"], g["K"], g["E"], g["src_depths"][s], g["src_K"][s],
                                                       g["src_E"][s], pix, dep)
        assert np.array_equal(m, g["masks"][s])
        assert np.array_equal(dr, g["reprojected"][s])
        assert np.array_equal(xs, g["x_src"][s], equal_nan=True) and np.array_equal(ys, g["y_src"][s], equal_nan=True)
    avg, pm, gm, fm, _ = fo.filter_view(g["ref_depth"], g["conf"], g["K"], g["E"], g["src_depths"], g["src_K"], g["src_E"],
                                        photomask=photo, geomask=int(geo), condmask_pixel=pix, condmask_depth=dep)
    assert np.array_equal(avg, g["depth_avg"]) and np.array_equal(gm, g["geo_mask"]) and np.array_equal(fm, g["final_mask"])
