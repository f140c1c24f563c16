#!/usr/bin/env python
"""Generate tests/golden/*.npz from the compiled reference (oracle/_ref).

Run in the build container only (needs /root/reference to have been compiled by
``python oracle/build_ref.py``).  Each fixture stores the case inputs and the
COO triplets / fint the reference's own per-element loop produced for it, so
the fixtures are self-contained on the GPU box.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loop  # noqa: E402
from tests import cases  # noqa: E402


def main():
    assert ref_loop.available(), "build oracle/_ref first: python oracle/build_ref.py"
    for name, case in cases.golden_cases().items():
        ref = ref_loop.run(case, state=True)
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
