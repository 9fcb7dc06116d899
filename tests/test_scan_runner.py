"""Scan-level runner (runner.ScanRunner / plan_scan): the host-side schedule on CPU, and on the GPU the guarantee that
sharing FeatureNet features across the reference views of a scan changes nothing: every depth / confidence map is
bit-identical to MVSNet.forward on the assembled per-view inputs (what the reference's eval loop feeds the model,
eval.py:326-360 with datasets/dataloader_eval.py:101-176)."""
import numpy as np
import pytest
import torch

from scene_3dreconstruction_mvsnet_b200.runner import plan_scan


def _pairs(n, nsrc):
    return [(i, [j for j in sorted(range(n), key=lambda j: (abs(j - i), j)) if j != i][:nsrc]) for i in range(n)]


def _check_plan(pairs, capacity):
    steps = plan_scan(pairs, capacity)
    resident = {}
    loads = 0
    for s, (ref, srcs) in zip(steps, pairs):
        assert s["views"] == [ref] + list(srcs)
        for img, slot in s["load"]:
            assert 0 <= slot < capacity
            for other in [i for i, sl in resident.items() if sl == slot]:
                assert other not in s["views"]      # eviction never touches a view of the current step
                del resident[other]
            resident[img] = slot
            loads += 1
        assert [resident[i] for i in s["views"]] == s["slots"]
        assert len(set(s["slots"])) == len(s["slots"])
    return loads


def test_plan_scan_loads_every_image_once_when_the_pool_is_large():
    pairs = _pairs(49, 4)                       # a DTU scan: 49 reference views x 5 views
    assert _check_plan(pairs, 64) == 49          # the reference extracts 49 x 5 = 245 feature maps for the same sweep


def test_plan_scan_small_pool_evicts_least_recently_used():
    pairs = _pairs(12, 4)
    loads = _check_plan(pairs, 5)                # exactly the views of one step fit
    assert 12 <= loads <= 12 * 5
    assert _check_plan(pairs, 7) <= loads


def test_plan_scan_ragged_and_errors():
    pairs = [(0, [1, 2]), (3, []), (2, [0]), (1, [3, 0, 2])]
    _check_plan(pairs, 4)
    with pytest.raises(ValueError):
        plan_scan([(0, [1, 2, 3])], 3)
    with pytest.raises(ValueError):
        plan_scan([(0, [1, 1])], 4)
    assert plan_scan([], 4) == []


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,capacity,streams", [(torch.float32, 8, 1), (torch.uint8, 3, 1), (torch.float32, 8, 2),
                                                     (torch.uint8, 3, 2), (torch.float32, 4, 3)])
def test_scan_runner_is_bit_identical_to_per_view_forward(weights, dtype, capacity, streams):
    from scene_3dreconstruction_mvsnet_b200 import synth
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet
    from scene_3dreconstruction_mvsnet_b200.runner import ScanRunner

    dev = "cuda:0"
    n, H, W, D = 6, 64, 96, 16
    g = torch.Generator().manual_seed(5)
    if dtype == torch.uint8:
        images = torch.randint(0, 256, (n, 3, H, W), dtype=torch.uint8, generator=g)
    else:
        images = torch.rand(n, 3, H, W, generator=g)
    cams = synth.make_cameras(n, H // 4, W // 4, focal=60.0, yaw=0.03, rolls=(4.0, -3.0, 6.0))
    projs = torch.from_numpy(cams)
    dv = 425.0 + 20.0 * torch.arange(D, dtype=torch.float32)
    pairs = _pairs(n, 2)                                    # 3 views per depth map; capacity 3 forces evictions
    model = MVSNet(refine=False, precision="bf16")
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    model = model.to(dev).eval()
    runner = ScanRunner(model, device=dev, pool_images=capacity, ring=2, depth=2, streams=streams)
    got = runner.run_scan([images[i] for i in range(n)], projs, dv, pairs)
    assert len(got) == n and runner.featurenet_images >= n
    if capacity >= n:
        assert runner.featurenet_images == n
    assert runner.h2d_bytes >= images.numel() * images.element_size()
    with torch.no_grad():
        for (ref, srcs), (depth, conf) in zip(pairs, got):
            ids = [ref] + srcs
            out = model(images[ids].unsqueeze(0).to(dev), projs[ids].unsqueeze(0).to(dev), dv.unsqueeze(0).to(dev))
            assert np.array_equal(out["depth"].cpu().numpy(), depth)
            assert np.array_equal(out["photometric_confidence"].cpu().numpy(), conf)
            assert float(np.std(depth)) > 0.0
    # a sink sees the same maps
    seen = {}
    runner.run_scan(images.numpy(), projs.numpy(), dv.numpy(), pairs, lambda k, d, c: seen.__setitem__(k, (d.copy(), c.copy())))
    assert sorted(seen) == list(range(n))
    assert all(np.array_equal(seen[k][0], got[k][0]) and np.array_equal(seen[k][1], got[k][1]) for k in range(n))


@pytest.mark.gpu
def test_pool_entry_rejects_bad_view_ids():
    from scene_3dreconstruction_mvsnet_b200 import ops
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet
    dev = "cuda:0"
    model = MVSNet(refine=False, precision="bf16").to(dev).eval()
    pool = torch.zeros((2, 16, 4, 24, 8), dtype=torch.float16, device=dev)
    proj = torch.eye(4, device=dev).repeat(1, 2, 1, 1)
    dv = (425.0 + torch.arange(16, dtype=torch.float32, device=dev)).unsqueeze(0)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="outside the pool"):
            model.forward_from_pool(pool, [0, 2], proj, dv)
        with pytest.raises(AssertionError):
            model.forward_from_pool(pool, [0], proj, dv)
    with pytest.raises(RuntimeError):
        ops.warp_variance_costreg_pool(pool.cpu(), [0, 1], proj, dv, model.cost_regularization.folded_params())


@pytest.mark.gpu
def test_depth_map_runner_cuda_graphs_match_eager(weights):
    """DepthMapRunner(graphs=True) replays one captured CUDA graph per staging slot: same maps, bit for bit, as the eager
    path, for inputs that change from view to view (the graph reads the slot's static device buffers)."""
    from scene_3dreconstruction_mvsnet_b200 import synth
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet
    from scene_3dreconstruction_mvsnet_b200.runner import DepthMapRunner
    dev = "cuda:0"
    model = MVSNet(refine=False, precision="bf16")
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    model = model.to(dev).eval()
    views = [synth.make_inputs(B=1, V=3, H=64, W=96, D=16, focal=90.0, interval_scale=8.0, seed=20 + i, yaw=0.01 * i)
             for i in range(11)]
    eager = DepthMapRunner(model, device=dev, depth=2).run_views(views)
    runner = DepthMapRunner(model, device=dev, depth=2, graphs=True)
    graphed = runner.run_views(views)
    assert all(s["graph"] is not None for s in runner._slots)
    for (d0, c0), (d1, c1) in zip(eager, graphed):
        assert np.array_equal(d0, d1) and np.array_equal(c0, c1)
    again = runner.run_views(views[::-1])
    for (d0, c0), (d1, c1) in zip(eager[::-1], again):
        assert np.array_equal(d0, d1) and np.array_equal(c0, c1)
    # weights change: the graphs (which hold the packed weights' addresses) are dropped and captured again
    with torch.no_grad():
        model.cost_regularization.conv0.conv.weight.mul_(1.5)
    want = DepthMapRunner(model, device=dev, depth=2).run_views(views)
    got = runner.run_views(views)
    assert not np.array_equal(want[0][0], eager[0][0])
    for (d0, c0), (d1, c1) in zip(want, got):
        assert np.array_equal(d0, d1) and np.array_equal(c0, c1)
