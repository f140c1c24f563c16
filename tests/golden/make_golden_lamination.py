#!/usr/bin/env python
"""Generate tests/golden/lamination.npz from the compiled reference (oracle/_ref): MatLamina invariants
(pyfe3d/shellprop.pyx:128-196), calc_lamination_parameters (:669), shellprop_from_LaminationParameters (:767) and
GradABDE.calc_LP_grad (:933) for random materials / thicknesses / lamination parameters, plus the lamination
parameters of a few random stacks.  Run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loop  # noqa: E402

FIELDS = ref_loop.SHELL_FIELDS
LP = ("xiA1", "xiA2", "xiA3", "xiA4", "xiB1", "xiB2", "xiB3", "xiB4", "xiD1", "xiD2", "xiD3", "xiD4", "xiE1", "xiE2")
INV = ("u1", "u2", "u3", "u4", "u5", "u6", "u7")


def main():
    assert ref_loop.available(), "build oracle/_ref first: python oracle/build_ref.py"
    ref_loop.load()
    from pyfe3d.shellprop import GradABDE, LaminationParameters, shellprop_from_LaminationParameters
    from pyfe3d.shellprop_utils import laminated_plate, read_laminaprop
    rng = np.random.default_rng(123)
    n = 12
    lam = np.stack([rng.uniform(50e9, 180e9, n), rng.uniform(5e9, 20e9, n), rng.uniform(0.2, 0.35, n),
                    rng.uniform(3e9, 7e9, n), rng.uniform(3e9, 7e9, n), rng.uniform(2e9, 5e9, n)], -1)
    thick = rng.uniform(0.5e-3, 8e-3, n)
    lps = rng.uniform(-1, 1, (n, 14))
    lps[0] = 0.                       # quasi-isotropic
    lps[1, 4:8] = 0.                  # symmetric
    inv = np.zeros((n, 7))
    props = np.zeros((n, 27))
    gA, gB, gD, gE = np.zeros((n, 6, 5)), np.zeros((n, 6, 5)), np.zeros((n, 6, 5)), np.zeros((n, 3, 3))
    for r in range(n):
        mat = read_laminaprop(tuple(lam[r]), 1500.)
        inv[r] = [getattr(mat, f) for f in INV]
        lp = LaminationParameters()
        for f, v in zip(LP, lps[r]):
            setattr(lp, f, v)
        p = shellprop_from_LaminationParameters(thick[r], mat, lp)
        props[r] = [getattr(p, f) for f in FIELDS]
        g = GradABDE()
        g.calc_LP_grad(thick[r], mat, lp)
        gA[r], gB[r], gD[r], gE[r] = np.asarray(g.gradAij), np.asarray(g.gradBij), np.asarray(g.gradDij), np.asarray(g.gradEij)
    # lamination parameters of real stacks (with and without offset) and the round trip through the LP form
    stacks = rng.uniform(-90, 90, (6, 5)).round(1)
    plyts = rng.uniform(0.1e-3, 1e-3, (6, 5))
    offs = rng.uniform(-1e-3, 1e-3, 6) * (np.arange(6) % 2)
    s_lp = np.zeros((6, 14))
    s_props = np.zeros((6, 27))
    for r in range(6):
        p = laminated_plate(stack=list(stacks[r]), plyts=list(plyts[r]), laminaprop=tuple(lam[r]), offset=float(offs[r]))
        q = p.calc_lamination_parameters()
        s_lp[r] = [getattr(q, f) for f in LP]
        s_props[r] = [getattr(p, f) for f in FIELDS]
    path = os.path.join(HERE, "lamination.npz")
    np.savez_compressed(path, lam=lam, thick=thick, lp=lps, inv=inv, props=props, gA=gA, gB=gB, gD=gD, gE=gE,
                        stacks=stacks, plyts=plyts, offs=offs, s_lp=s_lp, s_props=s_props)
    print("lamination.npz %.1f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
