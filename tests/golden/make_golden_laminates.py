#!/usr/bin/env python
"""Generate tests/golden/laminates.npz: ShellProp scalars of random laminates from the compiled reference
(oracle/_ref: laminated_plate pyfe3d/shellprop_utils.py:96, calc_constitutive_matrix shellprop.pyx:568, calc_scf :485).
Run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loop  # noqa: E402

FIELDS = ref_loop.SHELL_FIELDS


def main():
    assert ref_loop.available(), "build oracle/_ref first: python oracle/build_ref.py"
    ref_loop.load()
    from pyfe3d.shellprop_utils import laminated_plate
    rng = np.random.default_rng(77)
    flat = {}
    for g, nplies in enumerate((1, 3, 8)):
        nrows = 6
        stack = rng.uniform(-90, 90, (nrows, nplies)).round(1)
        plyts = rng.uniform(0.1e-3, 2e-3, (nrows, nplies))
        e1 = rng.uniform(50e9, 180e9, (nrows, nplies))
        e2 = rng.uniform(5e9, 20e9, (nrows, nplies))
        lam = np.stack([e1, e2, rng.uniform(0.2, 0.35, (nrows, nplies)), rng.uniform(3e9, 7e9, (nrows, nplies)),
                        rng.uniform(3e9, 7e9, (nrows, nplies)), rng.uniform(2e9, 5e9, (nrows, nplies))], -1)
        rhos = rng.uniform(1000., 8000., (nrows, nplies))
        offset = rng.uniform(-1e-3, 1e-3, nrows) * (g > 0)
        for scf in (0, 1):
            out = np.zeros((nrows, 27))
            for r in range(nrows):
                p = laminated_plate(stack=list(stack[r]), plyts=list(plyts[r]), laminaprops=[tuple(t) for t in lam[r]],
                                    rhos=list(rhos[r]), offset=float(offset[r]), calc_scf=bool(scf))
                out[r] = [getattr(p, f) for f in FIELDS]
            flat["g%d_scf%d_out" % (g, scf)] = out
        flat["g%d_stack" % g], flat["g%d_plyts" % g], flat["g%d_lam" % g] = stack, plyts, lam
        flat["g%d_rhos" % g], flat["g%d_offset" % g] = rhos, offset
    path = os.path.join(HERE, "laminates.npz")
    np.savez_compressed(path, **flat)
    print("laminates.npz %.1f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
