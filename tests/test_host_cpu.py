"""CPU-only tests: the C-ABI library loads and exports every declared symbol, host-side argument
checking, the nn.Module surface (state-dict compatibility, BN folding), view sharding."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from scene_3dreconstruction_mvsnet_b200 import _lib, ops
from scene_3dreconstruction_mvsnet_b200.models import MVSNet, mvsnet_loss
from scene_3dreconstruction_mvsnet_b200.models.module import fold_bn


def test_header_symbols_are_exported_and_bound():
    """Every function declared in include/mvsnet_b200.h is exported by the .so and bound in _lib.py."""
    hdr = open(os.path.join(ROOT, "include", "mvsnet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mvs_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.mvs_abi_version() == 2
    assert lib.mvs_arch() == b"sm_100a"


def test_workspace_queries_no_gpu():
    lib = _lib.load()
    n = lib.mvs_warp_variance_workspace_bytes(1, 5, 32, 288, 400)
    assert n >= 4 * 4 * 288 * 400 * 32
    assert lib.mvs_costreg_workspace_bytes(1, 192, 288, 400, 0) == int(23.75 * 192 * 288 * 400) * 4
    assert lib.mvs_costreg_workspace_bytes(1, 190, 288, 400, 0) == 0  # D % 8 != 0
    # argument validation happens before any CUDA call
    rc = lib.mvs_warp_variance_fwd(None, None, None, None, None, 1, 3, 32, 8, 8, 8, None)
    assert rc == -1 and b"null" in lib.mvs_last_error()


def test_state_dict_keys_match_reference(weights):
    m = MVSNet(refine=False)
    assert set(m.state_dict().keys()) == set(weights.keys())
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=True)
    assert sum(p.numel() for p in m.parameters()) == 338129
    # DataParallel-style prefix round trip (reference train.py:141, eval.py:315)
    sd = {"module." + k: v for k, v in m.state_dict().items()}
    torch.nn.DataParallel(MVSNet(refine=False)).load_state_dict(sd, strict=True)


def test_refine_true_is_rejected():
    with pytest.raises(NotImplementedError):
        MVSNet(refine=True)


def test_no_cpu_fallback():
    m = MVSNet(refine=False).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 3, 32, 32), torch.zeros(1, 3, 4, 4), torch.zeros(1, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.homo_warping(torch.zeros(1, 32, 8, 8), torch.zeros(1, 4, 4), torch.zeros(1, 4, 4), torch.zeros(1, 4))
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 3, 3, 32, 32), torch.zeros(1, 2, 4, 4), torch.zeros(1, 8))


def test_bn_folding_matches_unfolded(weights):
    m = MVSNet(refine=False)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    m.eval()
    layer = m.cost_regularization.conv2
    x = torch.randn(1, 16, 4, 6, 8)
    w, shift = layer.folded()
    y = torch.relu(torch.nn.functional.conv3d(x, w, None, 1, 1) + shift.view(1, -1, 1, 1, 1))
    assert torch.allclose(y, layer(x), atol=1e-5)
    seq = m.cost_regularization.conv9
    w, shift = fold_bn(seq[0].weight, seq[1], out_dim=1)
    x = torch.randn(1, 32, 2, 3, 4)
    y = torch.relu(torch.nn.functional.conv_transpose3d(x, w, None, stride=2, padding=1, output_padding=1) +
                   shift.view(1, -1, 1, 1, 1))
    assert torch.allclose(y, seq(x), atol=1e-5)
    # cache invalidation on in-place parameter update
    f1 = m.cost_regularization.folded_params()
    assert m.cost_regularization.folded_params() is f1
    with torch.no_grad():
        m.cost_regularization.conv0.conv.weight.mul_(2.0)
    assert m.cost_regularization.folded_params() is not f1


def test_mvsnet_loss_matches_definition():
    g = torch.Generator().manual_seed(0)
    est, gt = torch.randn(2, 5, 7, generator=g) * 3, torch.randn(2, 5, 7, generator=g)
    mask = (torch.rand(2, 5, 7, generator=g) > 0.4).float()
    d = (est - gt)[mask > 0.5].abs()
    ref = torch.where(d < 1, 0.5 * d * d, d - 0.5).mean()
    assert torch.allclose(mvsnet_loss(est, gt, mask), ref, atol=1e-6)


def test_oracle_is_not_imported_by_product():
    import subprocess
    import sys
    code = ("import sys; import scene_3dreconstruction_mvsnet_b200.models, scene_3dreconstruction_mvsnet_b200.ops, "
            "models; bad=[m for m in sys.modules if m.split('.')[0]=='oracle']; assert not bad, bad")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "scene_3dreconstruction_mvsnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("oracle/mvsnet_oracle.c:orc_compose_homography", ""), f


def test_folded_weight_cache_keys_on_cpu():
    """The folded-BN caches (host logic, no GPU): in-place updates through autograd's version counter are noticed,
    writes through .data need MVSNet.invalidate_folded(), a shallow module copy (what nn.DataParallel's replicate does
    to __dict__) does not inherit another module's tensor list, and ._apply (.to / .double) drops the caches."""
    import copy
    import torch
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet
    torch.manual_seed(0)
    m = MVSNet(refine=False).eval()
    cr = m.cost_regularization
    a = cr.folded_params()
    assert cr.folded_params() is a                                   # cached
    w0 = a[0][0].clone()
    with torch.no_grad():
        cr.conv0.conv.weight.mul_(2.0)                               # version bump
    b = cr.folded_params()
    assert b is not a and torch.allclose(b[0][0], 2.0 * w0)
    cr.conv0.conv.weight.data.mul_(0.5)                              # invisible to the key ...
    assert cr.folded_params() is b
    m.invalidate_folded()                                            # ... until told
    c = cr.folded_params()
    assert c is not b and torch.allclose(c[0][0], w0)
    # shallow copy with its own parameters: must not reuse the original's tensor list or folded weights
    r = copy.copy(cr)
    r._parameters = dict(cr._parameters)
    r._modules = {k: copy.deepcopy(v) for k, v in cr._modules.items()}
    with torch.no_grad():
        r.conv0.conv.weight.mul_(3.0)
    assert torch.allclose(r.folded_params()[0][0], 3.0 * w0) and torch.allclose(cr.folded_params()[0][0], w0)
    m.double()
    assert cr.folded_params()[0][0].dtype == torch.float64


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="the reference checkout only exists in the build container")
def test_public_surface_matches_the_reference_package():
    """Drop-in boundary: everything the reference's drivers import from `models` / `models.module`
    (train.py:15, eval.py:14, evalDTU.py:14; mvsnet.py:4) exists here under the same name with the reference's parameters in
    the reference's order (extra keyword parameters of ours come after them and have defaults)."""
    import inspect
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from make_golden import import_reference
    ref_mvsnet, ref_module = import_reference()
    import models as ours                      # the top-level shim a driver picks up through PYTHONPATH
    import models.module as ours_module

    def params(fn):
        return [p for p in inspect.signature(fn).parameters.values() if p.name != "self"]

    def check(ref_fn, our_fn, what):
        rp, op = params(ref_fn), params(our_fn)
        assert [p.name for p in op[:len(rp)]] == [p.name for p in rp], what
        for r, o in zip(rp, op):
            assert (r.default is inspect.Parameter.empty) == (o.default is inspect.Parameter.empty), (what, r.name)
            if r.default is not inspect.Parameter.empty:
                assert r.default == o.default, (what, r.name)
        assert all(p.default is not inspect.Parameter.empty for p in op[len(rp):]), what

    for name in ("MVSNet", "mvsnet_loss"):
        assert hasattr(ours, name), name
    check(ref_mvsnet.MVSNet.__init__, ours.MVSNet.__init__, "MVSNet.__init__")
    check(ref_mvsnet.MVSNet.forward, ours.MVSNet.forward, "MVSNet.forward")
    check(ref_mvsnet.mvsnet_loss, ours.mvsnet_loss, "mvsnet_loss")
    for name in ("homo_warping", "depth_regression"):
        check(getattr(ref_module, name), getattr(ours_module, name), name)
    for name in ("ConvBnReLU", "ConvBnReLU3D"):
        check(getattr(ref_module, name).__init__, getattr(ours_module, name).__init__, name)
    # same sub-module names, so checkpoints and code that reaches into the model (model.feature, ...) keep working
    ref_children = [n for n, _ in ref_mvsnet.MVSNet(refine=False).named_children()]
    our_children = [n for n, _ in ours.MVSNet(refine=False).named_children()]
    assert ref_children == our_children
