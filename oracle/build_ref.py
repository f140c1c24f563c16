#!/usr/bin/env python
"""Build the UNMODIFIED reference (saullocastro/pyfe3d) into ``oracle/_ref/pyfe3d``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyfe3d_b200/`` may import this.

The reference is Cython.  Its own build system (``/root/reference/setup.py:97-223``)
fails in this image at link time because it passes ``-fopenmp`` although no
``.pyx`` uses OpenMP, so we do not run it.  Instead each
``.pyx`` is cythonized *where it lies* under ``/root/reference`` (read-only;
only the generated ``.cpp`` goes to a scratch dir) and compiled with the flags
setuptools would have used (``-O2 -fno-strict-overflow -DNDEBUG``) into
``oracle/_ref/pyfe3d/*.so``.  No reference source is copied into the repo;
``oracle/_ref/`` is git-ignored but travels to the GPU box with gpurun.

The package ``__init__`` written below is ours (a minimal re-export), not a
copy of ``/root/reference/pyfe3d/__init__.py``.

Usage: ``python oracle/build_ref.py [--reference /root/reference] [--jobs N]``
"""
import argparse
import os
import subprocess
import sys
import sysconfig
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "pyfe3d")
MODULES = ["beamprop", "shellprop", "spring", "truss", "beamlr", "beamc",
           "tria3r", "quad4", "quad4r"]

INIT_PY = '''\
"""Reference pyfe3d, compiled by oracle/build_ref.py (generated file; test oracle only)."""
import ctypes as _ct
import importlib.util as _ilu
import os as _os
import sys as _sys
import numpy as _np
__version__ = "0.7.0+oracle"
from .quad4 import Quad4, Quad4Data, Quad4Probe
from .quad4r import Quad4R, Quad4RData, Quad4RProbe
from .tria3r import Tria3R, Tria3RData, Tria3RProbe
from .beamc import BeamC, BeamCData, BeamCProbe
from .beamlr import BeamLR, BeamLRData, BeamLRProbe
from .truss import Truss, TrussData, TrussProbe
from .spring import Spring, SpringData, SpringProbe
DOF = 6
INT = _np.int64 if _ct.sizeof(_ct.c_long) == 8 else _np.int32
DOUBLE = _np.float64
# The pure-python laminate helpers are not compiled code; when the read-only
# reference checkout is mounted (build container only) load them from there.
_p = "/root/reference/pyfe3d/shellprop_utils.py"
if _os.path.exists(_p):
    _spec = _ilu.spec_from_file_location(__name__ + ".shellprop_utils", _p)
    shellprop_utils = _ilu.module_from_spec(_spec)
    _sys.modules[__name__ + ".shellprop_utils"] = shellprop_utils
    _spec.loader.exec_module(shellprop_utils)
'''


def build_one(ref, tmp, mod, opt):
    pyx = os.path.join(ref, "pyfe3d", mod + ".pyx")
    cpp = os.path.join(tmp, "pyfe3d", mod + ".cpp")
    so = os.path.join(OUT, mod + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(so) and os.path.getmtime(so) > os.path.getmtime(pyx):
        return mod, "cached"
    subprocess.check_call([sys.executable, "-m", "cython", "--cplus", "-3",
                           "-I", ref, pyx, "-o", cpp])
    inc = sysconfig.get_paths()["include"]
    subprocess.check_call(["g++", opt, "-fPIC", "-shared", "-fno-strict-overflow",
                           "-DNDEBUG", "-w", "-I", inc, cpp, "-o", so])
    return mod, "built"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--jobs", type=int, default=min(8, os.cpu_count() or 1))
    ap.add_argument("--opt", default="-O2")
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(a.reference, "pyfe3d")):
        print("reference checkout not present at %s; nothing to build" % a.reference)
        return 0
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "pyfe3d"))
        with ThreadPoolExecutor(a.jobs) as ex:
            for mod, how in ex.map(lambda m: build_one(a.reference, tmp, m, a.opt), MODULES):
                print("oracle/_ref: %-10s %s" % (mod, how), flush=True)
    with open(os.path.join(OUT, "__init__.py"), "w") as f:
        f.write(INIT_PY)
    return 0


if __name__ == "__main__":
    sys.exit(main())
