"""Build the native pieces of pyfe3d_b200 in-tree (no JIT cache, no pip install):

  pyfe3d_b200/lib/libpyfe3d_b200.so   CUDA kernels + C ABI (include/pyfe3d_b200.h), sm_100a
  pyfe3d_b200/_cabi.*.so              thin Cython layer over that C ABI

``python -m pyfe3d_b200.build [--force]``; also called by ``__graft_entry__.build()``.
nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(PKG, "csrc", "_obj")
LIB = os.path.join(LIBDIR, "libpyfe3d_b200.so")
CU = ["quad_kernels.cu", "quad_fused.cu", "tria_kernels.cu", "tria_fused.cu", "props.cu", "line_kernels.cu", "assembly.cu", "solve.cu", "api.cu"]
HDRS = ["common.cuh", "shell.cuh", "pattern.hpp", os.path.join(ROOT, "include", "pyfe3d_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=True):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HDRS]

    def one(cu):
        src = os.path.join(CSRC, cu)
        obj = os.path.join(OBJDIR, cu[:-3] + ".o")
        if force or _newer(obj, [src] + hdrs):
            cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("PF3_EXTRA_NVCC_FLAGS", "").split() + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(len(CU)) as ex:
        objs = list(ex.map(one, CU))
    if force or _newer(LIB, objs):
        # link next to the target and rename: a reader (e.g. a gpurun snapshot) never sees a half-written library
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB + ".tmp"] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        os.replace(LIB + ".tmp", LIB)
    return LIB


def build_cabi(force=False, verbose=True):
    pyx = os.path.join(PKG, "_cabi.pyx")
    so = os.path.join(PKG, "_cabi" + sysconfig.get_config_var("EXT_SUFFIX"))
    hdr = os.path.join(ROOT, "include", "pyfe3d_b200.h")
    if not (force or _newer(so, [pyx, hdr])):
        return so
    c = os.path.join(OBJDIR, "_cabi.c")
    os.makedirs(OBJDIR, exist_ok=True)
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c])
    import numpy
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-w", "-I", sysconfig.get_paths()["include"],
           "-I", numpy.get_include(), "-I", os.path.join(ROOT, "include"), c, "-o", so + ".tmp",
           "-L", LIBDIR, "-lpyfe3d_b200", "-Wl,-rpath,$ORIGIN/lib"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.replace(so + ".tmp", so)
    return so


def build_all(force=False, verbose=True):
    build_lib(force, verbose)
    build_cabi(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
