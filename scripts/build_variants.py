#!/usr/bin/env python
"""Build experimental variants of the native library (quad_fused.cu compiled with extra -D flags) into
pyfe3d_b200/lib/variants/<name>/libpyfe3d_b200.so; scripts/gpu_variants.sh benches each one on the GPU box.
usage: python scripts/build_variants.py name1="-DX=1 -DY=2" name2="..." """
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfe3d_b200 import build as B  # noqa: E402

B.build_lib(verbose=False)
VARDIR = os.path.join(B.LIBDIR, "variants")


def one(arg):
    name, flags = arg.split("=", 1)
    d = os.path.join(VARDIR, name)
    os.makedirs(d, exist_ok=True)
    src = os.environ.get("VARIANT_SRC", "quad_fused.cu")
    obj = os.path.join(d, src[:-3] + ".o")
    subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + flags.split() + ["-c", os.path.join(B.CSRC, src), "-o", obj])
    objs = [obj if cu == src else os.path.join(B.OBJDIR, cu[:-3] + ".o") for cu in B.CU]
    subprocess.check_call([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o",
                           os.path.join(d, "libpyfe3d_b200.so")] + objs)
    os.remove(obj)
    r = subprocess.run(["cuobjdump", "-res-usage", os.path.join(d, "libpyfe3d_b200.so")], capture_output=True, text=True).stdout
    lines = r.splitlines()
    for i, ln in enumerate(lines):
        if "quad_fused_kernelILi0" in ln:
            print(name, lines[i + 1].strip()[:80])
    return name


with ThreadPoolExecutor(8) as ex:
    list(ex.map(one, sys.argv[1:]))
