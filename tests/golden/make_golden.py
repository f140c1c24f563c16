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


def sticky_sequence():
    """sticky_xmat.npz: the per-element objects keep m11..m22 across update_rotation_matrix calls whose xmat is null
    or parallel to the normal (quad4.pyx:485-488, 588, 598).  For element 6 of every <kind>_xmat_degenerate case
    (flat, in-plane material direction -> m != identity) the sequence  regular xmat -> parallel xmat -> null xmat
    is run on ONE reference object; m after each call and the KC0 values evaluated at the end are stored."""
    ref = ref_loop.load()
    flat = {}
    for kind in cases.SHELL_KINDS:
        case = cases.golden_cases()[kind + "_xmat_degenerate"]
        name = ref_loop._CLS[kind]
        data = getattr(ref, name + "Data")()
        el = getattr(ref, name)(getattr(ref, name + "Probe")())
        e = 6
        for a in range(case["conn"].shape[1]):
            setattr(el, "c%d" % (a + 1), int(6 * case["conn"][e, a]))
        x = np.ascontiguousarray(case["x"], float)
        ms = []
        for xm in (case["xmat"][e], (0., 0., 2.5), (0., 0., 0.), (0., 0., -1.)):
            el.update_rotation_matrix(x, float(xm[0]), float(xm[1]), float(xm[2]))
            ms.append([el.m11, el.m12, el.m21, el.m22])
        el.update_probe_xe(x)
        prop = ref_loop.make_props(kind, case["props"])[int(case["prop_id"][e])]
        n = data.KC0_SPARSE_SIZE
        r, c, v = np.zeros(n, ref.INT), np.zeros(n, ref.INT), np.zeros(n)
        el.update_KC0(r, c, v, prop)
        flat[kind + "_m"] = np.array(ms)
        flat[kind + "_KC0v"] = v
    path = os.path.join(HERE, "sticky_xmat.npz")
    np.savez_compressed(path, **flat)
    print("%-22s %7.1f kB" % ("sticky_xmat", os.path.getsize(path) / 1e3))


def main():
    """usage: make_golden.py [substring ...]: regenerate only the fixtures whose name contains a substring (all without
    arguments; existing fixtures are reproducible from their seeds)."""
    assert ref_loop.available(), "build oracle/_ref first: python oracle/build_ref.py"
    only = sys.argv[1:]
    if not only or any("sticky" in o for o in only):
        sticky_sequence()
    for name, case in cases.golden_cases().items():
        if only and not any(o in name for o in only):
            continue
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
