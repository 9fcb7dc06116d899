"""Generates tests/golden/io_* with the UNMODIFIED reference's datasets/data_io.py (run in the build container, where
/root/reference exists):   python tests/make_golden_io.py
  io_gray.pfm, io_gray1.pfm, io_color.pfm, io_scaled.pfm   written by the reference's save_pfm
  io_rgb.png, io_gray.png                                  synthetic source images (written here with PIL)
  io_cases.npz                                             the arrays that were saved + the reference's
                                                           read_rescale_crop_img outputs for the two images
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, "/root/reference")
from datasets.data_io import read_pfm, read_rescale_crop_img, save_pfm  # noqa: E402  (the reference's own)
from PIL import Image  # noqa: E402

rng = np.random.default_rng(7)
gray = (rng.standard_normal((7, 5)) * 300).astype(np.float32)
gray1 = rng.random((4, 6, 1)).astype(np.float32)
color = rng.random((3, 4, 3)).astype(np.float32)
save_pfm(os.path.join(GOLD, "io_gray.pfm"), gray)
save_pfm(os.path.join(GOLD, "io_gray1.pfm"), gray1)
save_pfm(os.path.join(GOLD, "io_color.pfm"), color)
save_pfm(os.path.join(GOLD, "io_scaled.pfm"), gray, scale=2.5)
back, scale = read_pfm(os.path.join(GOLD, "io_scaled.pfm"))
assert np.array_equal(back, gray) and scale == 2.5

rgb = rng.integers(0, 256, (100, 140, 3), dtype=np.uint8)
g8 = rng.integers(0, 256, (100, 140), dtype=np.uint8)
Image.fromarray(rgb).save(os.path.join(GOLD, "io_rgb.png"))
Image.fromarray(g8).save(os.path.join(GOLD, "io_gray.png"))
K = np.array([[120.0, 0, 70.0], [0, 121.0, 50.0], [0, 0, 1]], np.float32)
out = {"gray": gray, "gray1": gray1, "color": color, "K": K}
for tag, res in (("rgb", (64, 96)), ("gray", (64, 96)), ("rgb_b", (96, 128))):
    k = K.copy()
    img, k2 = read_rescale_crop_img(os.path.join(GOLD, "io_%s.png" % tag.split("_")[0]), k, img_res=res)
    out["img_" + tag] = img
    out["K_" + tag] = k2
    out["res_" + tag] = np.array(res)
np.savez_compressed(os.path.join(GOLD, "io_cases.npz"), **out)
print({k: v.shape for k, v in out.items()})
