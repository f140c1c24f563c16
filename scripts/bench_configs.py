#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (device-resident, values only, CUDA events).
Not the headline metric (bench.py is); prints one JSON line per config."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfe3d_b200 import meshes  # noqa: E402
from pyfe3d_b200.batch import AssemblyPlan, ElementBatch  # noqa: E402
from tests import util  # noqa: E402


def timeit(fn, steps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def run(name, case, mats, fused, scale=1.0):
    b = util.batch_from_case(case)
    nn = case["ndof"] // 6
    ne = case["conn"].shape[0]
    kw = {}
    for m in mats:
        if m == "KGs":
            kw["KG_given_stress"] = case.get("stress", (0., 0., 1.))
        elif m.startswith("M"):
            kw["M"] = True
            kw["mtype"] = int(m[1])
        else:
            kw[m] = True
    names = sorted(set("KG" if m == "KGs" else ("M" if m.startswith("M") else m) for m in mats))
    mtype = kw.get("mtype", 0)
    plans = {m: AssemblyPlan(m, nn, [b], mtype=mtype) for m in names}
    if fused:
        coo, csr = plans["KC0"].evaluate_assemble(**kw)
        ms = timeit(lambda: plans["KC0"].evaluate_assemble(coo=coo, csr=csr, **kw))
    else:
        coo = b.evaluate(indices=False, **kw)
        csr = {m: torch.empty(plans[m].nnz, dtype=torch.float64, device=b.device) for m in names}

        def step():
            b.evaluate(indices=False, out=coo, **kw)
            for m in names:
                plans[m].assemble(coo[m].v, out=csr[m])
        ms = timeit(step)
        parts = {"eval": timeit(lambda: b.evaluate(indices=False, out=coo, **kw))}
        for m in names:
            parts["assemble_" + m] = timeit(lambda m=m: plans[m].assemble(coo[m].v, out=csr[m]))
        print(json.dumps({"config": name, "kernel_ms": parts}))
    bytes_el = sum(b.sizes[m] * 8 for m in names)
    nnz = sum(plans[m].nnz for m in names)
    alg = ne * bytes_el + nnz * 8
    print(json.dumps({"config": name, "kind": case["kind"], "elements": ne, "matrices": list(mats),
                      "path": "fused" if fused else "two-pass", "ms_per_step": ms, "elements_per_s": ne / ms * 1e3,
                      "algorithmic_GBps": alg / ms / 1e6, "frac_of_6538.9": alg / ms / 1e6 / 6538.9}))
    del plans, coo, csr, b
    torch.cuda.empty_cache()


def run_spmv(side):
    """y = KC0 x on the assembled matrix of the north-star mesh, per-entry CSR against the
    plan's block structure."""
    from pyfe3d_b200.batch import spmv
    case = meshes.plate_quad4(side, side)
    b = util.batch_from_case(case)
    nn = case["ndof"] // 6
    plan = AssemblyPlan("KC0", nn, [b])
    _, csr = plan.evaluate_assemble(KC0=True, write_coo=False)
    vals = csr["KC0"]
    x = torch.randn(6 * nn, dtype=torch.float64, device=vals.device)
    y = torch.empty(6 * nn, dtype=torch.float64, device=vals.device)
    free = (torch.rand(6 * nn, device=vals.device) > 0.05).to(torch.uint8)
    nnz, nblk = plan.nnz, plan._plan.nblocks
    ms_plan = timeit(lambda: plan.spmv(vals, x, out=y), steps=10)
    ms_mask = timeit(lambda: plan.spmv(vals, x, free=free, out=y), steps=10)
    alg_plan = nnz * 8 + nblk * 8 + 2 * 6 * nn * 8
    print(json.dumps({"config": "spmv KC0 %dx%d Quad4 (plan block structure)" % (side, side), "nnz": nnz,
                      "ms": ms_plan, "ms_masked": ms_mask, "algorithmic_GBps": alg_plan / ms_plan / 1e6,
                      "frac_of_6538.9": alg_plan / ms_plan / 1e6 / 6538.9}))
    indptr, indices = plan.pattern()
    ms_csr = timeit(lambda: spmv(indptr, indices, vals, x, out=y), steps=10)
    alg_csr = nnz * 16 + 2 * 6 * nn * 8 + 6 * nn * 8
    print(json.dumps({"config": "spmv KC0 %dx%d Quad4 (int64 CSR)" % (side, side), "nnz": nnz, "ms": ms_csr,
                      "algorithmic_GBps": alg_csr / ms_csr / 1e6, "frac_of_6538.9": alg_csr / ms_csr / 1e6 / 6538.9}))


def run_aero(side):
    """the three piston-theory matrices of the north-star mesh in one launch (values only)."""
    case = meshes.plate_quad4(side, side)
    b = util.batch_from_case(case)
    coo = b.evaluate_aero(KA_beta=True, KA_gamma=True, CA=True, indices=False)
    ms = timeit(lambda: b.evaluate_aero(KA_beta=True, KA_gamma=True, CA=True, indices=False, out=coo))
    ne = case["conn"].shape[0]
    alg = ne * (3 * 144 * 8 + 32 + 24)
    print(json.dumps({"config": "aero KA_beta+KA_gamma+CA %dx%d Quad4" % (side, side), "elements": ne, "ms": ms,
                      "elements_per_s": ne / ms * 1e3, "algorithmic_GBps": alg / ms / 1e6,
                      "frac_of_6538.9": alg / ms / 1e6 / 6538.9}))


if __name__ == "__main__":
    if "--aero" in sys.argv:
        run_aero(200 if "--small" in sys.argv else 2000)
        sys.exit(0)
    if "--spmv" in sys.argv:
        run_spmv(200 if "--small" in sys.argv else 2000)
        sys.exit(0)
    small = "--small" in sys.argv
    f = 0.1 if small else 1.0
    if "--config4" in sys.argv:
        run("config4 Tria3R distorted plate 4M, KC0+M(mtype1)", meshes.plate_tria3r(int(1415 * f ** 0.5), int(1415 * f ** 0.5)),
            ("KC0", "M1"), True)
        sys.exit(0)
    run("config2 BeamC arc 100k, KC0+M", meshes.arc_beamc(int(100001 * f)), ("KC0", "M0"), False)
    run("config3 Quad4R cylinder 1M, KC0+KG_given_stress", meshes.cylinder_quad4r(int(1760 * f ** 0.5), int(571 * f ** 0.5)),
        ("KC0", "KGs"), True)
    run("config3 (two-pass)", meshes.cylinder_quad4r(int(1760 * f ** 0.5), int(571 * f ** 0.5)), ("KC0", "KGs"), False)
    run("config4 Tria3R distorted plate 4M, KC0+M(mtype1)", meshes.plate_tria3r(int(1415 * f ** 0.5), int(1415 * f ** 0.5)),
        ("KC0", "M1"), True)
    run("config4 (two-pass)", meshes.plate_tria3r(int(1415 * f ** 0.5), int(1415 * f ** 0.5)), ("KC0", "M1"), False)
