"""CPU restatement (numpy, float64 like the reference) of the geometric-consistency filter that follows the depth path.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and tests/make_golden_fusion.py), never by the product.

Follows the reference (olivier-2018/scene_3Dreconstruction_MVSNet):
  eval.py:508-560  reproject_with_depth          (project ref pixels into the source view, sample its depth, project back)
  eval.py:564-585  check_geometric_consistency    (|p_reproj - p| < condmask_pixel and |d_reproj - d| / d < condmask_depth)
  eval.py:660-703  filter_depth, per reference view: photo mask, per-source masks, averaged depth, geometric / final mask
`cv2.remap(..., INTER_LINEAR)` (eval.py:540) is restated too: float maps are quantised to 1/32 pixel
(cvRound(x * 32), integer part >> 5, 5-bit fractions), weights (1-fy)(1-fx), ... formed in float32, the four taps
accumulated left to right in float32, taps outside the image read the constant border value 0; non-finite coordinates
saturate to -32768.  Checked bit-exact against cv2.remap 4.13 in the build container (tests/make_golden_fusion.py).
Pinned against the unmodified reference functions through tests/golden/fusion_*.npz.
"""
import numpy as np


def remap_bilinear(src, map_x, map_y):
    """cv2.remap(src, map_x, map_y, interpolation=cv2.INTER_LINEAR) for float32 src / maps (border constant 0)."""
    h, w = src.shape
    f32 = np.float32
    with np.errstate(invalid="ignore", over="ignore"):
        sx = np.rint((map_x.astype(f32) * f32(32)).astype(np.float64))
        sy = np.rint((map_y.astype(f32) * f32(32)).astype(np.float64))
    bad = ~np.isfinite(sx) | ~np.isfinite(sy) | (np.abs(sx) > 2 ** 31 - 1) | (np.abs(sy) > 2 ** 31 - 1)
    sx = np.where(bad, -2 ** 31, sx).astype(np.int64)
    sy = np.where(bad, -2 ** 31, sy).astype(np.int64)
    ax = (sx & 31).astype(f32) / f32(32)
    ay = (sy & 31).astype(f32) / f32(32)
    ix = np.clip(sx >> 5, -32768, 32767)
    iy = np.clip(sy >> 5, -32768, 32767)
    w00, w01 = (f32(1) - ay) * (f32(1) - ax), (f32(1) - ay) * ax
    w10, w11 = ay * (f32(1) - ax), ay * ax

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        return np.where(ok, src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], f32(0)).astype(f32)

    v = tap(iy, ix) * w00
    v = (v + tap(iy, ix + 1) * w01).astype(f32)
    v = (v + tap(iy + 1, ix) * w10).astype(f32)
    v = (v + tap(iy + 1, ix + 1) * w11).astype(f32)
    return v


def reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    """eval.py:508-560.  Returns depth_reprojected, x_reprojected, y_reprojected, x_src, y_src (all float32 [h,w])."""
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    x_ref, y_ref = x_ref.reshape([-1]), y_ref.reshape([-1])
    ones = np.ones_like(x_ref)
    with np.errstate(divide="ignore", invalid="ignore"):
        xyz_ref = np.matmul(np.linalg.inv(intrinsics_ref), np.vstack((x_ref, y_ref, ones)) * depth_ref.reshape([-1]))
        xyz_src = np.matmul(np.matmul(extrinsics_src, np.linalg.inv(extrinsics_ref)), np.vstack((xyz_ref, ones)))[:3]
        k_xyz_src = np.matmul(intrinsics_src, xyz_src)
        xy_src = k_xyz_src[:2] / k_xyz_src[2:3]
        x_src = xy_src[0].reshape([height, width]).astype(np.float32)
        y_src = xy_src[1].reshape([height, width]).astype(np.float32)
        sampled = remap_bilinear(depth_src.astype(np.float32), x_src, y_src)
        xyz_src = np.matmul(np.linalg.inv(intrinsics_src), np.vstack((xy_src, ones)) * sampled.reshape([-1]))
        xyz_rep = np.matmul(np.matmul(extrinsics_ref, np.linalg.inv(extrinsics_src)), np.vstack((xyz_src, ones)))[:3]
        depth_rep = xyz_rep[2].reshape([height, width]).astype(np.float32)
        k_xyz_rep = np.matmul(intrinsics_ref, xyz_rep)
        xy_rep = k_xyz_rep[:2] / k_xyz_rep[2:3]
    x_rep = xy_rep[0].reshape([height, width]).astype(np.float32)
    y_rep = xy_rep[1].reshape([height, width]).astype(np.float32)
    return depth_rep, x_rep, y_rep, x_src, y_src


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                condmask_pixel=1.0, condmask_depth=0.01):
    """eval.py:564-585.  Returns mask (bool), depth_reprojected (float32, 0 where masked out), x_src, y_src."""
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    depth_rep, x_rep, y_rep, x_src, y_src = reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src,
                                                                 intrinsics_src, extrinsics_src)
    with np.errstate(divide="ignore", invalid="ignore"):
        dist = np.sqrt((x_rep - x_ref) ** 2 + (y_rep - y_ref) ** 2)
        depth_diff = np.abs(depth_rep - depth_ref)
        relative = depth_diff / depth_ref
        mask = np.logical_and(dist < condmask_pixel, relative < condmask_depth)
    depth_rep = depth_rep.copy()
    depth_rep[~mask] = 0
    return mask, depth_rep, x_src, y_src


def filter_view(ref_depth, confidence, ref_intrinsics, ref_extrinsics, src_depths, src_intrinsics, src_extrinsics,
                photomask=0.8, geomask=3, condmask_pixel=1.0, condmask_depth=0.01):
    """eval.py:660-703 for one reference view and its source views.
    Returns depth_est_averaged (float64), photo_mask, geo_mask, final_mask (bool), geo_mask_sum (int32)."""
    photo_mask = confidence > photomask
    reprojected = []
    geo_mask_sum = 0
    for d, k, e in zip(src_depths, src_intrinsics, src_extrinsics):
        m, dr, _, _ = check_geometric_consistency(ref_depth, ref_intrinsics, ref_extrinsics, d, k, e, condmask_pixel,
                                                  condmask_depth)
        geo_mask_sum = geo_mask_sum + m.astype(np.int32)
        reprojected.append(dr)
    depth_avg = (sum(reprojected) + ref_depth) / (geo_mask_sum + 1)
    geo_mask = geo_mask_sum >= geomask
    final_mask = np.logical_and(photo_mask, geo_mask)
    return depth_avg, photo_mask, geo_mask, final_mask, np.asarray(geo_mask_sum, dtype=np.int32)
