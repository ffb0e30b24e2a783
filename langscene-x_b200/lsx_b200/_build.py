"""Builds liblsx_b200.so (the C-ABI library of include/lsx_rasterizer.h) in-tree with nvcc for sm_100a.

The library depends only on the CUDA runtime — no torch headers — so it compiles in well under a
minute and the same binary is loaded by ctypes from any Python (or bound from C/C++ directly).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)                      # langscene-x_b200/
REPO = os.path.dirname(ROOT)
CSRC = os.path.join(ROOT, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(PKG_DIR, "liblsx_b200.so")

SOURCES = ["api.cu", "preprocess.cu", "sort.cu", "binning.cu", "cull.cu", "render_fwd.cu", "render_bwd.cu", "stats.cu", "knn.cu", "depth_normal.cu", "image_loss.cu", "arena_adam.cu", "gaussian_head.cu", "densify.cu", "pose.cu", "cls3d.cu", "microbench.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-Wno-attributes",
    # NOTE: no -use_fast_math: radii / tile rects / depth keys / alpha thresholds are bit-exact contracts
]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(REPO, "include", "lsx_rasterizer.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    if not _stale(obj, [src] + _headers()):
        return obj
    cmd = ["nvcc", *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout[-3000:]}\n{r.stderr[-6000:]}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build_library(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if force or _stale(LIB_PATH, objs):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", LIB_PATH, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))


ROUTE_B_SRC = os.path.join(ROOT, "integration", "rasterize_points_lsx.cpp")
ROUTE_B_LIB = os.path.join(ROOT, "integration", "lsx_route_b.so")


def build_route_b(force=False):
    """INTEGRATION.md Route B as a binary: the reference's own C++ glue (pybind `_C` functions with the reference's signatures)
    re-bound on the C ABI, compiled with g++ against torch's headers and linked with liblsx_b200.so.  It proves that the stub
    the integration guide shows compiles and binds; tests/test_route_b.py runs it next to the ctypes route on the GPU."""
    import sysconfig
    import torch
    from torch.utils import cpp_extension as ce
    build_library()
    if not force and not _stale(ROUTE_B_LIB, [ROUTE_B_SRC, LIB_PATH, os.path.join(REPO, "include", "lsx_rasterizer.h")]):
        return ROUTE_B_LIB
    inc = ce.include_paths("cuda") + [sysconfig.get_paths()["include"], os.path.join(REPO, "include")]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", ROUTE_B_SRC, "-o", ROUTE_B_LIB,
           *[f"-I{p}" for p in inc], f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-DTORCH_EXTENSION_NAME=lsx_route_b",
           f"-L{PKG_DIR}", "-l:liblsx_b200.so", f"-Wl,-rpath,{PKG_DIR}", "-Wl,-rpath,$ORIGIN/../lsx_b200"]
    for p in ce.library_paths("cuda"):
        cmd += [f"-L{p}", f"-Wl,-rpath,{p}"]
    cmd += ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"Route B glue failed to build:\n{r.stderr[-6000:]}")
    return ROUTE_B_LIB
