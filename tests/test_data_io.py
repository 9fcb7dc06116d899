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


@pytest.mark.gpu
def test_eval_loop_from_disk_to_pfm(weights, tmp_path):
    """The reference's eval loop end to end on a synthetic on-disk scan (eval.py:326-396 with
    datasets/dataloader_eval.py:101-176): PNG images + cam files -> data_io loaders -> ScanRunner -> PfmWriter; the
    PFMs read back equal MVSNet.forward on the float32 images the reference's loader would have produced."""
    import torch
    from PIL import Image
    from scene_3dreconstruction_mvsnet_b200 import synth
    from scene_3dreconstruction_mvsnet_b200.models import MVSNet
    from scene_3dreconstruction_mvsnet_b200.runner import ScanRunner

    n, H, W, D = 5, 64, 96, 16
    rng = np.random.default_rng(3)
    cams = synth.make_cameras(n, H // 4, W // 4, focal=60.0, yaw=0.02).astype(np.float64)
    Kq = np.array([[60.0, 0, W / 8.0], [0, 60.0, H / 8.0], [0, 0, 1]])
    for i in range(n):
        Image.fromarray(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).save(tmp_path / ("%08d.png" % i))
        E = np.eye(4)
        E[:3, :4] = np.linalg.inv(Kq) @ cams[i][:3, :4]
        K_full = Kq.copy()
        K_full[:2] *= 4.0                                          # cam files hold full-resolution intrinsics
        lines = ["extrinsic"] + [" ".join("%.9g" % v for v in r) for r in E] + ["", "intrinsic"] + \
                [" ".join("%.9g" % v for v in r) for r in K_full] + ["", "425.0 20.0"]
        (tmp_path / ("%08d_cam.txt" % i)).write_text("\n".join(lines) + "\n")

    images_u8, images_f32, projs = [], [], []
    for i in range(n):
        K, E, dmin, dint = data_io.read_cam_file(str(tmp_path / ("%08d_cam.txt" % i)))
        img8, K2 = data_io.read_rescale_crop_img(str(tmp_path / ("%08d.png" % i)), K.copy(), img_res=(H, W), as_uint8=True)
        imgf, _ = data_io.read_rescale_crop_img(str(tmp_path / ("%08d.png" % i)), K.copy(), img_res=(H, W))
        K2[:2] /= 4.0                                              # dataloader_eval.py: intrinsics at feature resolution
        P = E.copy()
        P[:3, :4] = K2 @ E[:3, :4]
        images_u8.append(img8)
        images_f32.append(np.ascontiguousarray(imgf.transpose(2, 0, 1)))
        projs.append(P)
    projs = np.stack(projs).astype(np.float32)
    dv = (dmin + dint * np.arange(D)).astype(np.float32)
    pairs = [(i, [j for j in sorted(range(n), key=lambda j: (abs(j - i), j)) if j != i][:2]) for i in range(n)]

    model = MVSNet(refine=False, precision="bf16")
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    model = model.to("cuda:0").eval()
    out_dir = tmp_path / "out"
    path_of = lambda k: (str(out_dir / "depth_est" / ("%08d.pfm" % k)), str(out_dir / "confidence" / ("%08d.pfm" % k)))
    with data_io.PfmWriter(threads=2) as wr:
        ScanRunner(model, device="cuda:0", pool_images=8).run_scan(images_u8, projs, dv, pairs, wr.sink(path_of))
    assert wr.files_written == 2 * n
    with torch.no_grad():
        for k, (ref, srcs) in enumerate(pairs):
            ids = [ref] + srcs
            imgs = torch.from_numpy(np.stack([images_f32[j] for j in ids])).unsqueeze(0).cuda()
            out = model(imgs, torch.from_numpy(projs[ids]).unsqueeze(0).cuda(), torch.from_numpy(dv).unsqueeze(0).cuda())
            depth, scale = data_io.read_pfm(path_of(k)[0])
            conf, _ = data_io.read_pfm(path_of(k)[1])
            assert scale == 1.0 and depth.shape == (H // 4, W // 4)
            assert np.array_equal(depth, out["depth"][0].cpu().numpy())
            assert np.array_equal(conf, out["photometric_confidence"][0].cpu().numpy())
            assert float(depth.std()) > 0
