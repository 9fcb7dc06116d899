"""Builds libmvsnet_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m scene_3dreconstruction_mvsnet_b200.build [--force]

The library has no torch dependency: it exports the C ABI declared in include/mvsnet_b200.h and is
loaded with ctypes (scene_3dreconstruction_mvsnet_b200/_lib.py).  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmvsnet_b200.so")
SOURCES = ["api.cu", "warp_variance.cu", "warp_variance_win.cu", "depth_tail.cu", "conv3d_fp32.cu", "conv2d_fp32.cu", "conv3d_tc.cu", "fusion.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libmvsnet_b200.so cannot be built")


def _deps(src):
    return [os.path.join(CSRC, src), os.path.join(CSRC, "common.cuh"),
            os.path.join(os.path.dirname(HERE), "include", "mvsnet_b200.h")] + \
        [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc"))]


def _compile(nvcc, src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + ["-Xptxas", "-v"] * bool(verbose) + \
        ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build_library(force=False, verbose=False, report=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    todo = []
    for src in SOURCES:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        newest = max(os.path.getmtime(p) for p in _deps(src))
        if force or verbose or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
            todo.append(src)
    logs = {}
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=len(todo)) as ex:
            for src, (obj, log) in zip(todo, ex.map(lambda s: _compile(nvcc, s, verbose), todo)):
                logs[src] = log
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if todo or not os.path.exists(LIB):
        cmd = [nvcc] + ARCH + ["-shared", "-cudart", "shared", "-o", LIB] + objs + \
            ["-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if report:
        print("libmvsnet_b200.so: compiled %s; reused up-to-date objects of %s" % (
            ", ".join(todo) or "nothing", ", ".join(s for s in SOURCES if s not in todo) or "nothing"))
    if verbose:
        for src, log in logs.items():
            print("==== %s\n%s" % (src, log))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
