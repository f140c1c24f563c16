import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ensure_built():
    """The native library is built in-tree and git-ignored: build it once if a fresh checkout lacks it."""
    import glob
    import subprocess
    lib = os.path.join(ROOT, "pyfe3d_b200", "lib", "libpyfe3d_b200.so")
    ext = glob.glob(os.path.join(ROOT, "pyfe3d_b200", "_cabi*.so"))
    if not (os.path.exists(lib) and ext):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "pyfe3d_b200", "build.py")])


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    _ensure_built()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
