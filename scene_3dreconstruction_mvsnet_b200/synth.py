"""Seeded synthetic inputs for the depth-inference path.

There is no dataset and no checkpoint in the build environment, so every test and every bench
run uses images, cameras and depth hypotheses generated here.  The shapes and camera model follow
the reference's data contract (datasets/dataloader_eval.py:101-176 of the reference):
imgs [B,V,3,H,W] in [0,1], proj_matrices [B,V,4,4] with K (already /4) @ [R|t] in rows 0-2 and
0 0 0 1 in row 3, depth_values [B,D] = depth_min + interval * arange(D), all fp32.
"""
import math

import numpy as np
import torch

# name -> (V, H, W, D, focal at 1/4 res, interval_scale)   (BASELINE.json configs)
CONFIGS = {
    "c1_3view_512x640": (3, 512, 640, 192, 361.5, 1.06),
    "c2_dtu_5view_1152x1600": (5, 1152, 1600, 192, 723.0, 1.06),
    "c3_bin_4view_512x640": (4, 512, 640, 192, 361.5, 1.33),
}

# rotated-camera variants (DTU cameras are not rectified): name -> (base config, yaw towards the scene centre in
# rad per 60 mm of baseline, roll of views 1.. in degrees).  Non-axis-aligned source footprints exercise the window
# planner of the fused warp kernel (segment halving, wider windows) that the rectified set-up never triggers.
ROTATED = {
    "c2_dtu_5view_1152x1600_rot": ("c2_dtu_5view_1152x1600", 0.1, (-8.0, 5.0, -10.0, 7.0)),
    "c1_3view_512x640_rot": ("c1_3view_512x640", 0.1, (-8.0, 5.0)),
}

_BASELINES_MM = (0.0, -60.0, 60.0, -120.0, 120.0, -180.0, 180.0, -240.0)


def make_cameras(V, h, w, focal, yaw=0.0, dtype=np.float32, converge=0.0, rolls=None):
    """[V,4,4] projection matrices at feature resolution (h, w) for V cameras on an x-baseline.
    converge: yaw (rad per 60 mm of baseline) turning every source camera towards the scene in front of the reference
    camera; rolls: rotation about the optical axis (degrees) of views 1.."""
    K = np.array([[focal, 0.0, w / 2.0], [0.0, focal, h / 2.0], [0.0, 0.0, 1.0]], np.float64)
    out = np.zeros((V, 4, 4), np.float64)
    for v in range(V):
        a = yaw * (v % 3 - 1) if v else 0.0
        base = _BASELINES_MM[v % len(_BASELINES_MM)]
        a += converge * base / 60.0
        R = np.array([[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]])
        if rolls is not None and v:
            r = math.radians(rolls[(v - 1) % len(rolls)])
            R = np.array([[math.cos(r), -math.sin(r), 0.0], [math.sin(r), math.cos(r), 0.0], [0.0, 0.0, 1.0]]) @ R
        E = np.eye(4)
        E[:3, :3] = R
        E[0, 3] = _BASELINES_MM[v % len(_BASELINES_MM)]
        E[1, 3] = 7.0 * (v % 2)
        P = E.copy()
        P[:3, :4] = K @ E[:3, :4]
        out[v] = P
    return out.astype(dtype)


def make_inputs(B=1, V=3, H=512, W=640, D=192, focal=361.5, interval_scale=1.06, depth_min=425.0, yaw=0.0,
                seed=0, gray=False, converge=0.0, rolls=None):
    """Returns CPU tensors (imgs, proj_matrices, depth_values)."""
    g = torch.Generator().manual_seed(seed)
    if gray:  # datasets/data_io.py:149-150 replicates the single channel
        imgs = torch.rand(B, V, 1, H, W, generator=g).expand(B, V, 3, H, W).contiguous()
    else:
        imgs = torch.rand(B, V, 3, H, W, generator=g)
    cams = make_cameras(V, H // 4, W // 4, focal, yaw, converge=converge, rolls=rolls)
    proj = torch.from_numpy(np.broadcast_to(cams, (B, V, 4, 4)).copy())
    interval = 2.5 * interval_scale
    dv = depth_min + interval * torch.arange(D, dtype=torch.float32)
    depth_values = dv.unsqueeze(0).repeat(B, 1)
    for b in range(1, B):  # per-sample depth ranges differ, like a real batch
        depth_values[b] += 3.0 * b
    return imgs, proj, depth_values


def config_of(name):
    """(V, H, W, D, focal, interval_scale) of a named workload, rotated variants included."""
    return CONFIGS[ROTATED[name][0] if name in ROTATED else name]


def make_named(name, B=1, seed=0):
    converge, rolls = 0.0, None
    if name in ROTATED:
        name, converge, rolls = ROTATED[name]
    V, H, W, D, focal, itv = CONFIGS[name]
    return make_inputs(B=B, V=V, H=H, W=W, D=D, focal=focal, interval_scale=itv, seed=seed,
                       gray=name.startswith("c3"), converge=converge, rolls=rolls)


def make_features(B, V, C, h, w, seed=0):
    """Smooth-ish random feature maps [B,V,C,h,w] (for kernel-level tests and the kernel bench)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, V, C, h, w, generator=g)
