"""CPU: the C-ABI shared library loads and exports every symbol include/pyfe3d_b200.h declares;
static tables agree with the reference's XData classes; compute calls fail loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyfe3d_b200", "lib", "libpyfe3d_b200.so")
HDR = os.path.join(ROOT, "include", "pyfe3d_b200.h")


def _declared():
    txt = open(HDR).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pf3_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build first: python pyfe3d_b200/build.py"
    lib = ctypes.CDLL(LIB)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert lib.pf3_version() >= 100


def test_sparse_sizes_match_reference_data_classes():
    lib = ctypes.CDLL(LIB)
    # (KC0, KG, M)_SPARSE_SIZE of quad4.pyx:174-180, quad4r.pyx:140-146, tria3r.pyx:148-151, beamc.pyx:43-46,
    # beamlr.pyx:43-46, truss.pyx:45-47, spring.pyx:41-43
    want = {0: (576, 144, 480), 1: (576, 144, 480), 2: (324, 81, 270), 3: (144, 144, 144), 4: (144, 36, 144),
            5: (72, 0, 144), 6: (72, 0, 0)}
    for kind, sizes in want.items():
        assert tuple(lib.pf3_sparse_size(kind, m) for m in range(3)) == sizes
    # KA_BETA / KA_GAMMA / CA_SPARSE_SIZE (quad4.pyx:150-152): the two quads only
    for kind in range(7):
        assert tuple(lib.pf3_sparse_size(kind, m) for m in (3, 4, 5)) == ((144,) * 3 if kind < 2 else (0,) * 3)
    assert lib.pf3_written_size(0, 2, 2) == 288 and lib.pf3_written_size(2, 2, 2) == 162
    assert lib.pf3_written_size(3, 2, 1) == 36 and lib.pf3_written_size(4, 1, 0) == 36
    assert [lib.pf3_num_nodes(k) for k in range(7)] == [4, 4, 3, 2, 2, 2, 2]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import pyfe3d_b200 as pf
    from pyfe3d_b200.batch import ElementBatch
    with pytest.raises(RuntimeError):
        ElementBatch("quad4", np.zeros((1, 4), np.int64), np.zeros(12), np.zeros((1, 32)))
    q = pf.Quad4(pf.Quad4Probe())
    q.c1, q.c2, q.c3, q.c4 = 0, 6, 12, 18
    with pytest.raises(RuntimeError):
        q.update_probe_xe(np.zeros(12))
    lib = ctypes.CDLL(LIB)
    ctx = ctypes.c_void_p()
    assert lib.pf3_create(0, ctypes.byref(ctx)) == -2     # PF3_E_NO_DEVICE
    lib.pf3_error_string.restype = ctypes.c_char_p
    assert b"no CPU fallback" in lib.pf3_error_string(-2)
