"""Geometric-consistency filter of the depth maps on the GPU: the mirror of the reference's
`check_geometric_consistency` (eval.py:564-585) and of the per-view part of `filter_depth` (eval.py:660-703).

Same argument meaning and numpy in / numpy out as the reference functions, so the reference's fusion loop can call
these instead; the work is one CUDA kernel (csrc/fusion.cu, `mvs_filter_depth`) per reference view instead of ~60 numpy
passes and one cv2.remap per source view.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _dev(a, dtype, device):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype)).to(device)


def _mats(m, shape):
    m = np.ascontiguousarray(m, dtype=np.float64)
    if m.shape != shape:
        raise RuntimeError("camera matrix stack has shape %s, expected %s" % (m.shape, shape))
    return m


def filter_view(ref_depth, confidence, ref_intrinsics, ref_extrinsics, src_depths, src_intrinsics, src_extrinsics,
                photomask=0.8, geomask=3, condmask_pixel=1.0, condmask_depth=0.01, device="cuda:0", details=False):
    """One reference view against its S source views (eval.py:660-703 with the thresholds of eval.py:45-49).
    ref_depth, confidence [h,w]; src_depths [S,h,w]; intrinsics 3x3 / extrinsics 4x4 per view (numpy).
    Returns a dict of numpy arrays: depth_est_averaged (float64), photo_mask, geo_mask, final_mask (bool), geo_mask_sum
    (int32) and, with details=True, per source view depth_reprojected / mask / x2d_src / y2d_src."""
    if not torch.cuda.is_available():
        raise RuntimeError("filter_view runs on CUDA devices only; there is no CPU fallback")
    ref_depth = np.asarray(ref_depth)
    if ref_depth.ndim != 2:
        raise RuntimeError("ref_depth must be [h, w], got %s" % (ref_depth.shape,))
    h, w = ref_depth.shape
    src_depths = np.asarray(src_depths, dtype=np.float32).reshape(-1, h, w)
    S = src_depths.shape[0]
    kr, er = _mats(ref_intrinsics, (3, 3)), _mats(ref_extrinsics, (4, 4))
    ks, es = _mats(np.asarray(src_intrinsics).reshape(-1, 3, 3), (S, 3, 3)), _mats(np.asarray(src_extrinsics).reshape(-1, 4, 4), (S, 4, 4))
    dev = torch.device(device)
    d_ref = _dev(ref_depth, np.float32, dev)
    d_conf = _dev(confidence, np.float32, dev) if confidence is not None else None
    d_src = _dev(src_depths, np.float32, dev) if S else None
    avg = torch.empty((h, w), dtype=torch.float64, device=dev)
    gsum = torch.empty((h, w), dtype=torch.int32, device=dev)
    pm, gm, fm = (torch.empty((h, w), dtype=torch.uint8, device=dev) for _ in range(3))
    rep = torch.empty((S, h, w), dtype=torch.float32, device=dev) if details and S else None
    sm = torch.empty((S, h, w), dtype=torch.uint8, device=dev) if details and S else None
    xy = torch.empty((S, 2, h, w), dtype=torch.float32, device=dev) if details and S else None
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    hp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    with torch.cuda.device(dev):
        rc = _lib.load().mvs_filter_depth(p(d_ref), p(d_conf), hp(kr), hp(er), p(d_src), hp(ks) if S else None,
                                          hp(es) if S else None, S, h, w, float(condmask_pixel), float(condmask_depth),
                                          int(geomask), float(photomask), p(avg), p(gsum), p(pm), p(gm), p(fm), p(rep), p(sm),
                                          p(xy), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "mvs_filter_depth")
    out = {"depth_est_averaged": avg.cpu().numpy(), "photo_mask": pm.cpu().numpy().astype(bool),
           "geo_mask": gm.cpu().numpy().astype(bool), "final_mask": fm.cpu().numpy().astype(bool),
           "geo_mask_sum": gsum.cpu().numpy()}
    if details and S:
        out.update(depth_reprojected=rep.cpu().numpy(), mask=sm.cpu().numpy().astype(bool), x2d_src=xy[:, 0].cpu().numpy(),
                   y2d_src=xy[:, 1].cpu().numpy())
    return out


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                condmask_pixel=1.0, condmask_depth=0.01, device="cuda:0"):
    """Drop-in for eval.py:564 (the two thresholds are `args.condmask_pixel` / `args.condmask_depth` there).
    Returns mask, depth_reprojected, x2d_src, y2d_src."""
    r = filter_view(depth_ref, None, intrinsics_ref, extrinsics_ref, np.asarray(depth_src)[None], np.asarray(intrinsics_src)[None],
                    np.asarray(extrinsics_src)[None], condmask_pixel=condmask_pixel, condmask_depth=condmask_depth,
                    device=device, details=True)
    return r["mask"][0], r["depth_reprojected"][0], r["x2d_src"][0], r["y2d_src"][0]
