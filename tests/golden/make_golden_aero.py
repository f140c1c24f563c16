#!/usr/bin/env python
"""Generate tests/golden/aero_*.npz: piston-theory matrices KA_beta / KA_gamma / CA of Quad4 and Quad4R from the
compiled reference (oracle/_ref; update_KA_beta quad4.pyx:9491, update_KA_gamma :10312, update_CA :11115 and the
Quad4R twins quad4r.pyx:12789, :13605, :14403), driven like tests/test_quad4r_piston_theory.py:84-124.
Run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loop  # noqa: E402
from tests import cases  # noqa: E402

WHAT = ("KA_beta", "KA_gamma", "CA")


def aero_cases():
    out = {}
    for k in ("quad4", "quad4r"):
        out["aero_%s_mesh" % k] = cases.shell_mesh(k, 6, 5, seed=31)
        out["aero_%s_soup" % k] = cases.shell_soup(k, 20, seed=32)
    return out


def main():
    assert ref_loop.available(), "build oracle/_ref first: python oracle/build_ref.py"
    for name, case in aero_cases().items():
        ref = ref_loop.run(case, what=WHAT, state=True)
        flat = {}
        for k, v in case.items():
            if v is None or isinstance(v, str):
                continue
            flat["in_" + k] = np.asarray(v)
        flat["in_kind"] = np.array(case["kind"])
        for k, v in ref.items():
            if isinstance(v, list):
                flat["ref_%s_r" % k], flat["ref_%s_c" % k], flat["ref_%s_v" % k] = v
            else:
                flat["ref_" + k] = v
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **flat)
        print("%-22s %7.1f kB" % (name, os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
