"""I/O around the depth-inference path: the reference's file formats, kept byte for
byte, moved off the critical path.

  save_pfm / read_pfm          datasets/data_io.py:8-73 of the reference: `Pf\\n{w} {h}\\n-1.000000\\n` + bottom-up
                               little-endian float32 rows (`PF` and 3 channels for colour images)
  read_rescale_crop_img        datasets/data_io.py:76-154: PIL decode, bilinear down-scale, centre crop to a multiple of
                               32, intrinsics adjusted; here it can also return the image as uint8 [3,H,W] -- the /255
                               then happens on the device (mvs_featurenet_tc_fwd_u8) after a 4x smaller upload
  read_cam_file                datasets/dataloader_eval.py:56-71
  PfmWriter                    the reference writes both PFMs of a view synchronously inside its eval loop
                               (eval.py:387-392), i.e. the GPU idles while the host formats and writes 0.9 MB per view;
                               here a small pool of writer threads takes copies of the runner's pinned output slots

Host code only: no CUDA, no torch.
"""
import math
import os
import queue
import re
import sys
import threading

import numpy as np


def _pfm_header(image, scale):
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        color = True
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        color = False
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    endian = image.dtype.byteorder
    if endian == "<" or (endian == "=" and sys.byteorder == "little"):
        scale = -scale
    return ("PF\n" if color else "Pf\n") + "{} {}\n".format(image.shape[1], image.shape[0]) + ("%f\n" % scale)


def pfm_bytes(image, scale=1):
    """The exact byte string the reference's save_pfm writes for `image` (float32 [H,W], [H,W,1] or [H,W,3])."""
    image = np.asarray(image)
    return _pfm_header(image, scale).encode("utf-8") + np.ascontiguousarray(image[::-1]).tobytes()


def save_pfm(filename, image, scale=1):
    with open(filename, "wb") as f:
        f.write(pfm_bytes(image, scale))


def read_pfm(filename):
    """-> (data [H,W] or [H,W,3] float32 in the file's byte order, top row first; scale)."""
    with open(filename, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header == "PF":
            color = True
        elif header == "Pf":
            color = False
        else:
            raise Exception("Not a PFM file.")
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not m:
            raise Exception("Malformed PFM header.")
        width, height = map(int, m.groups())
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        data = np.frombuffer(f.read(), dtype=endian + "f4")
    shape = (height, width, 3) if color else (height, width)
    if data.size != height * width * (3 if color else 1):
        raise Exception("PFM payload does not match its header.")
    return np.flipud(data.reshape(shape)), abs(scale)


def read_cam_file(filename, interval_scale=1.0):
    """-> (intrinsics 3x3 f32, extrinsics 4x4 f32, depth_min, depth_interval * interval_scale)."""
    with open(filename) as f:
        lines = [line.rstrip() for line in f.readlines()]
    extrinsics = np.array(" ".join(lines[1:5]).split(), dtype=np.float32).reshape(4, 4)
    intrinsics = np.array(" ".join(lines[7:10]).split(), dtype=np.float32).reshape(3, 3)
    depth_min = float(lines[11].split()[0])
    depth_interval = float(lines[11].split()[1]) * interval_scale
    return intrinsics, extrinsics, depth_min, depth_interval


def read_rescale_crop_img(img_fname, intrinsics, img_res=(512, 640), as_uint8=False):
    """Decode, down-scale (PIL bilinear, the larger of the two scale factors), centre-crop to the target size or to a
    multiple of 32, adjust the intrinsics IN PLACE like the reference, and return (image, intrinsics).
    as_uint8=False: float32 [H,W,3] in [0,1] (the reference's output).  as_uint8=True: uint8 [3,H,W], channel-planar,
    ready for a 4x smaller upload; uint8 / 255 on the device is bit-identical to the float32 form."""
    from PIL import Image
    base = 32
    img = Image.open(img_fname)
    w_src, h_src = img.size
    h_target, w_target = img_res
    h_scale, w_scale = float(h_target) / h_src, float(w_target) / w_src
    if h_scale > 1 or w_scale > 1:
        raise ValueError("img_res %s exceeds the image size (%d, %d): images are only ever reduced" % (img_res, h_src, w_src))
    resize_scale = max(h_scale, w_scale)
    img = img.resize(size=(int(w_src * resize_scale), int(h_src * resize_scale)), resample=Image.BILINEAR)
    w_res, h_res = img.size
    intrinsics[:2, :] *= resize_scale
    final_h = h_target if h_res > h_target else int(math.floor(h_target / base) * base)
    final_w = w_target if w_res > w_target else int(math.floor(w_target / base) * base)
    start_h = int(math.floor((h_res - final_h) / 2))
    start_w = int(math.floor((w_res - final_w) / 2))
    img = img.crop((start_w, start_h, start_w + final_w, start_h + final_h))
    intrinsics[0, -1] -= start_w
    intrinsics[1, -1] -= start_h
    if as_uint8:
        a = np.array(img, dtype=np.uint8)
        if a.ndim == 2:
            a = np.stack((a, a, a), axis=2)
        return np.ascontiguousarray(a.transpose(2, 0, 1)), intrinsics
    np_img = np.array(img, dtype=np.float32) / 255.
    if np_img.ndim == 2:
        np_img = np.dstack((np_img, np_img, np_img))
    return np_img, intrinsics


class PfmWriter:
    """Background PFM writer: submit(path, array) copies the array (a runner's pinned slot is reused two views later)
    and returns at once; `threads` workers format and write.  close() (or leaving the `with` block) waits for every
    file and re-raises the first error."""

    def __init__(self, threads=2, max_pending=16, makedirs=True):
        self._q = queue.Queue(maxsize=max_pending)
        self._err = None
        self._makedirs = makedirs
        self.files_written = 0
        self.bytes_written = 0
        self._lock = threading.Lock()
        self._workers = [threading.Thread(target=self._run, daemon=True) for _ in range(threads)]
        for w in self._workers:
            w.start()

    def _run(self):
        while True:
            item = self._q.get()
            try:
                if item is None:
                    return
                path, arr, scale = item
                if self._makedirs:
                    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
                data = pfm_bytes(arr, scale)
                with open(path, "wb") as f:
                    f.write(data)
                with self._lock:
                    self.files_written += 1
                    self.bytes_written += len(data)
            except Exception as e:  # noqa: BLE001 -- reported by close()
                with self._lock:
                    if self._err is None:
                        self._err = e
            finally:
                self._q.task_done()

    def submit(self, path, array, scale=1):
        if self._err is not None:
            raise self._err
        self._q.put((path, np.array(array, dtype=np.float32, copy=True), scale))

    def sink(self, path_of):
        """A `sink(index, depth, conf)` for DepthMapRunner.run_views / ScanRunner.run_scan: path_of(index) ->
        (depth_path, confidence_path), as in eval.py:376-392."""
        def _sink(index, depth, conf):
            dpath, cpath = path_of(index)
            self.submit(dpath, depth[0] if depth.ndim == 3 else depth)
            self.submit(cpath, conf[0] if conf.ndim == 3 else conf)
        return _sink

    def close(self):
        self._q.join()
        for _ in self._workers:
            self._q.put(None)
        for w in self._workers:
            w.join()
        if self._err is not None:
            raise self._err

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
