"""I/O around the path against files and arrays produced by the unmodified reference's
datasets/data_io.py (tests/make_golden_io.py): the PFM writer must be byte-exact, the image loader value-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from scene_3dreconstruction_mvsnet_b200 import data_io


def _bytes(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="module")
def cases():
    return load_golden("io_cases.npz")


def test_pfm_writer_is_byte_exact(cases, tmp_path):
    assert data_io.pfm_bytes(cases["gray"]) == _bytes("io_gray.pfm")
    assert data_io.pfm_bytes(cases["gray1"]) == _bytes("io_gray1.pfm")
    assert data_io.pfm_bytes(cases["color"]) == _bytes("io_color.pfm")
    assert data_io.pfm_bytes(cases["gray"], scale=2.5) == _bytes("io_scaled.pfm")
    assert _bytes("io_gray.pfm").startswith(b"Pf\n5 7\n-1.000000\n")          # the wire format SURVEY 8(f) names
    p = tmp_path / "x.pfm"
    data_io.save_pfm(str(p), cases["gray"][:, ::-1])                          # non-contiguous input
    assert p.read_bytes() == data_io.pfm_bytes(np.ascontiguousarray(cases["gray"][:, ::-1]))


def test_pfm_reader_and_errors(cases, tmp_path):
    for name, key, scale in (("io_gray.pfm", "gray", 1.0), ("io_color.pfm", "color", 1.0), ("io_scaled.pfm", "gray", 2.5)):
        data, s = data_io.read_pfm(os.path.join(GOLDEN, name))
        assert s == scale and np.array_equal(data, cases[key])
    data, _ = data_io.read_pfm(os.path.join(GOLDEN, "io_gray1.pfm"))
    assert np.array_equal(data, cases["gray1"][:, :, 0])
    with pytest.raises(Exception, match="float32"):
        data_io.pfm_bytes(cases["gray"].astype(np.float64))
    with pytest.raises(Exception, match="dimensions"):
        data_io.pfm_bytes(np.zeros((2, 2, 2), np.float32))
    bad = tmp_path / "bad.pfm"
    bad.write_bytes(b"P5\n1 1\n-1.0\n\0\0\0\0")
    with pytest.raises(Exception, match="Not a PFM"):
        data_io.read_pfm(str(bad))


@pytest.mark.parametrize("tag", ["rgb", "gray", "rgb_b"])
def test_image_loader_matches_reference(cases, tag):
    res = tuple(int(v) for v in cases["res_" + tag])
    fname = os.path.join(GOLDEN, "io_%s.png" % tag.split("_")[0])
    k = cases["K"].copy()
    img, k2 = data_io.read_rescale_crop_img(fname, k, img_res=res)
    assert img.dtype == np.float32 and np.array_equal(img, cases["img_" + tag])
    assert np.array_equal(k2, cases["K_" + tag]) and k2 is k                   # adjusted in place, like the reference
    u8, _ = data_io.read_rescale_crop_img(fname, cases["K"].copy(), img_res=res, as_uint8=True)
    assert u8.dtype == np.uint8 and u8.shape == (3,) + img.shape[:2]
    assert np.array_equal(u8.transpose(1, 2, 0).astype(np.float32) / np.float32(255.0), img)
    with pytest.raises(ValueError):
        data_io.read_rescale_crop_img(fname, cases["K"].copy(), img_res=(512, 640))


def test_cam_file_roundtrip(tmp_path):
    E = np.arange(16, dtype=np.float32).reshape(4, 4) / 7
    K = np.array([[2892.33, 0, 823.2], [0, 2883.18, 619.07], [0, 0, 1]], np.float32)
    lines = ["extrinsic"] + [" ".join(repr(float(v)) for v in r) for r in E] + ["", "intrinsic"] + \
            [" ".join(repr(float(v)) for v in r) for r in K] + ["", "425.0 2.5"]
    p = tmp_path / "00000000_cam.txt"
    p.write_text("\n".join(lines) + "\n")
    k, e, dmin, dint = data_io.read_cam_file(str(p), interval_scale=1.06)
    assert np.array_equal(k, K) and np.array_equal(e, E) and dmin == 425.0 and dint == 2.5 * 1.06


def test_async_writer(cases, tmp_path):
    paths = {}

    def path_of(i):
        paths[i] = (str(tmp_path / "depth_est" / ("%08d.pfm" % i)), str(tmp_path / "confidence" / ("%08d.pfm" % i)))
        return paths[i]

    slot = np.empty((1, 7, 5), np.float32)                                   # a reused "pinned slot"
    with data_io.PfmWriter(threads=3, max_pending=4) as wr:
        sink = wr.sink(path_of)
        for i in range(12):
            slot[0] = cases["gray"] + i
            sink(i, slot, slot * 0.5)
    assert wr.files_written == 24
    for i in range(12):
        d, _ = data_io.read_pfm(paths[i][0])
        c, _ = data_io.read_pfm(paths[i][1])
        assert np.array_equal(d, cases["gray"] + i) and np.array_equal(c, (cases["gray"] + i) * 0.5)
    wr2 = data_io.PfmWriter(threads=1, makedirs=False)
    wr2.submit(str(tmp_path / "missing_dir" / "a.pfm"), cases["gray"])
    with pytest.raises(OSError):
        wr2.close()
