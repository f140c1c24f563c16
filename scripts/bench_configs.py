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


def run_cg(side, iters=64, host_side=400):
    """one iteration of the native Jacobi-CG (pf3_plan_cg) on KC0 of the north-star mesh against its
    roofline (block SpMV: 8 B per nonzero + 8 B per 6x6 block; the three vector kernels: 13 vector passes), and scipy's
    cg on the host cores on a smaller plate of the same mesh (per-iteration time scaled by the nonzero count)."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import cg
    from pyfe3d_b200.solve import plan_cg_native
    case, free, f, normal = meshes.static_case(side)
    b = meshes.batch_from_case(case)
    nn = case["ndof"] // 6
    plan = AssemblyPlan("KC0", nn, [b])
    _, csr = plan.evaluate_assemble(KC0=True, write_coo=False)
    vals = csr["KC0"]
    free_t = torch.as_tensor(free.astype(np.uint8)).cuda()
    ft = torch.as_tensor(f).cuda()
    plan_cg_native(plan, vals, ft, free=free_t, rtol=0., maxiter=8, graph=False)          # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, it, status, res, bn = plan_cg_native(plan, vals, ft, free=free_t, rtol=0., maxiter=iters, graph=False, check_every=iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = 6 * nn
    alg = plan.nnz * 8 + plan._plan.nblocks * 8 + 2 * n * 8 + 13 * n * 8
    ms_it = dt / it * 1e3
    out = {"config": "native CG iteration, KC0 of %dx%d Quad4 (%d dofs, %d nnz), clamped edges" % (side, side, n, plan.nnz),
           "iterations_timed": it, "ms_per_iteration": ms_it, "algorithmic_bytes_per_iteration": alg,
           "achieved_GBps": alg / ms_it / 1e6, "frac_of_6538.9": alg / ms_it / 1e6 / 6538.9}
    nnz_dev = plan.nnz
    del plan, vals, csr, b
    torch.cuda.empty_cache()
    # host: scipy cg on the diagonally scaled Kuu of a host_side x host_side plate of the same mesh
    case, free, f, normal = meshes.static_case(host_side)
    b = meshes.batch_from_case(case)
    plan = AssemblyPlan("KC0", case["ndof"] // 6, [b])
    _, csr = plan.evaluate_assemble(KC0=True, write_coo=False)
    K = plan.to_scipy(csr["KC0"]).tocsr()
    bu = free.astype(bool)
    Kuu = K[bu, :][:, bu]
    dis = 1.0 / np.sqrt(np.maximum(Kuu.diagonal(), 1e-30))
    D = sp.diags(dis)
    Ks = (D @ Kuu @ D).tocsr()
    fs = D @ f[bu]
    t0 = time.perf_counter()
    cg(Ks, fs, rtol=0., atol=0., maxiter=iters)
    dth = (time.perf_counter() - t0) / iters * 1e3
    out["scipy_host_ms_per_iteration"] = dth
    out["scipy_host_nnz"] = int(Ks.nnz)
    out["scipy_host_ms_per_iteration_scaled_to_device_nnz"] = dth * nnz_dev / Ks.nnz
    out["note"] = "host figure: scipy cg (1 thread SpMV) on the %dx%d plate; last field scales it by the nonzero count" % (host_side, host_side)
    print(json.dumps(out))


def run_dropin(n=2000):
    """Cost of the PER-ELEMENT drop-in classes (import pyfe3d_b200 as pyfe3d): the reference's own loop
    (tests/test_quad4_static_point_load.py:53-78) over n Quad4 elements, one C-ABI round trip per method call, against
    the batched API on the same mesh."""
    import pyfe3d_b200 as pf
    from pyfe3d_b200.shellprop_utils import isotropic_plate
    side = int(n ** 0.5)
    case = meshes.plate_quad4(side, side, with_u=True)
    x = np.ascontiguousarray(case["x"], float)
    u = np.ascontiguousarray(case["u"], float)
    conn = case["conn"]
    ne = conn.shape[0]
    prop = isotropic_plate(thickness=0.005, E=200e9, nu=0.3, calc_scf=True, rho=7800.)
    data, probe = pf.Quad4Data(), pf.Quad4Probe()
    KC0r = np.zeros(data.KC0_SPARSE_SIZE * ne, pf.INT); KC0c = np.zeros_like(KC0r); KC0v = np.zeros(KC0r.size)
    KGr = np.zeros(data.KG_SPARSE_SIZE * ne, pf.INT); KGc = np.zeros_like(KGr); KGv = np.zeros(KGr.size)
    Mr = np.zeros(data.M_SPARSE_SIZE * ne, pf.INT); Mc = np.zeros_like(Mr); Mv = np.zeros(Mr.size)
    t0 = time.perf_counter()
    for e in range(ne):
        q = pf.Quad4(probe)
        q.n1, q.n2, q.n3, q.n4 = [int(t) for t in conn[e]]
        q.c1, q.c2, q.c3, q.c4 = [int(6 * t) for t in conn[e]]
        q.init_k_KC0, q.init_k_KG, q.init_k_M = e * data.KC0_SPARSE_SIZE, e * data.KG_SPARSE_SIZE, e * data.M_SPARSE_SIZE
        q.update_rotation_matrix(x)
        q.update_probe_xe(x)
        q.update_KC0(KC0r, KC0c, KC0v, prop)
        q.update_probe_ue(u)
        q.update_KG(KGr, KGc, KGv, prop)
        q.update_M(Mr, Mc, Mv, prop)
    dt = time.perf_counter() - t0
    b = ElementBatch("quad4", conn, x, prop, u=u)
    b.evaluate(KC0=True, KG=True, M=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    coo = b.evaluate(KC0=True, KG=True, M=True)
    v = coo["KC0"].v.cpu().numpy()
    dtb = time.perf_counter() - t1
    err = float(np.abs(v - KC0v).max() / np.abs(KC0v).max())
    print(json.dumps({"config": "per-element drop-in loop, %d Quad4 (6 method calls per element)" % ne, "seconds": dt,
                      "us_per_method_call": dt / (6 * ne) * 1e6, "elements_per_s": ne / dt,
                      "batched_api_seconds_same_mesh_incl_indices_and_d2h": dtb, "KC0_max_rel_diff_loop_vs_batch": err}))


def run_mixed(side, nstiff=64):
    """Config 5 (per-GPU share): Quad4 skin + BeamC stiffeners in ONE matrix per KC0 / KG / M, plus update_fint.
    The groups have different masks for KG and M, so this is the two-pass path: one evaluation launch per group and
    the per-node gather assembly over the union pattern."""
    skin, beams = meshes.stiffened_panel(side, side, nstiff=nstiff)
    bs = [util.batch_from_case(skin), util.batch_from_case(beams)]
    nn = skin["ndof"] // 6
    ne = sum(b.ne for b in bs)
    names = ("KC0", "KG", "M")
    plans = {m: AssemblyPlan(m, nn, bs) for m in names}
    coos = [b.evaluate(KC0=True, KG=True, M=True, indices=False) for b in bs]
    vals = {m: torch.cat([c[m].v for c in coos]) for m in names}
    offs = {m: [0, coos[0][m].v.numel()] for m in names}
    views = [{m: type(coos[g][m])(None, None, vals[m][offs[m][g]:offs[m][g] + coos[g][m].v.numel()], 6 * nn)
              for m in names} for g in range(2)]
    csr = {m: torch.empty(plans[m].nnz, dtype=torch.float64, device=bs[0].device) for m in names}
    fint = torch.zeros(6 * nn, dtype=torch.float64, device=bs[0].device)

    def step():
        for g, b in enumerate(bs):
            b.evaluate(KC0=True, KG=True, M=True, indices=False, out=views[g])
        for m in names:
            plans[m].assemble(vals[m], out=csr[m])
        plans["KC0"].update_fint(fint)

    ms = timeit(step)
    coo_f = {m: type(coos[0][m])(None, None, vals[m], 6 * nn) for m in names}

    def fstep():
        plans["KC0"].evaluate_assemble(KC0=True, KG=True, M=True, coo=coo_f, csr=csr)
        plans["KC0"].update_fint(fint)

    ms_f = timeit(fstep)
    from pyfe3d_b200 import _cabi
    from pyfe3d_b200.batch import _ptr, context
    pk = plans["KC0"]
    what = _cabi.KC0 | _cabi.KG | _cabi.M
    cc = lambda m: _cabi.Coo(0, 0, _ptr(vals[m]), 0, 0)
    context(bs[0].device)
    fparts = {"fused_quad_group": timeit(lambda: pk._plan.eval_assemble_group(bs[0].cabi_batch(), 0, what, cc("KC0"), cc("KG"), cc("M"),
                                                                             _ptr(csr["KC0"]), _ptr(csr["KG"]), _ptr(csr["M"]))),
              "beam_eval": timeit(lambda: bs[1].evaluate(KC0=True, KG=True, M=True, indices=False, out=views[1]))}
    for m in names:
        fparts["add_" + m] = timeit(lambda m=m: plans[m]._plan.assemble_add(_ptr(vals[m]), _ptr(csr[m]), 0))
    print(json.dumps({"config": "config5 fused quad share + beams added", "ms_per_step": ms_f,
                      "elements_per_s": ne / ms_f * 1e3, "kernel_ms": fparts}))
    parts = {"eval": timeit(lambda: [b.evaluate(KC0=True, KG=True, M=True, indices=False, out=views[g]) for g, b in enumerate(bs)])}
    for m in names:
        parts["assemble_" + m] = timeit(lambda m=m: plans[m].assemble(vals[m], out=csr[m]))
    parts["fint"] = timeit(lambda: plans["KC0"].update_fint(fint))
    print(json.dumps({"config": "config5 share", "kernel_ms": parts}))
    alg = sum(sum(b.ne * b.sizes[m] for b in bs) * 8 + plans[m].nnz * 8 for m in names)
    print(json.dumps({"config": "config5 stiffened panel %dx%d Quad4 + %d BeamC, KC0+KG+M+fint" % (side, side, bs[1].ne),
                      "kind": "quad4+beamc", "elements": ne, "matrices": list(names), "path": "two-pass", "ms_per_step": ms,
                      "elements_per_s": ne / ms * 1e3, "algorithmic_GBps": alg / ms / 1e6,
                      "frac_of_6538.9": alg / ms / 1e6 / 6538.9}))


def run_kinds(n):
    """every element kind, KC0 + KG + M (whatever the kind has) + fint, n elements, values only."""
    from tests import cases
    for kind in ("quad4", "quad4r", "tria3r", "beamc", "beamlr", "truss", "spring"):
        if kind in ("quad4", "quad4r"):
            side = int(n ** 0.5)
            case = meshes.plate_quad4(side, side, kind=kind)
        elif kind == "tria3r":
            side = int((n / 2) ** 0.5)
            case = meshes.plate_tria3r(side, side)
        else:
            case = cases.line_chain(kind, n, seed=1)
        b = util.batch_from_case(case)
        kw = dict(KC0=True, KG=b.sizes["KG"] > 0, M=b.sizes["M"] > 0, indices=False)
        coo = b.evaluate(**kw)
        ms = timeit(lambda: b.evaluate(out=coo, **kw))
        bytes_el = sum(b.sizes[m] for m in ("KC0", "KG", "M")) * 8
        fint = torch.zeros(case["ndof"], dtype=torch.float64, device=b.device)
        plan = AssemblyPlan("KC0", case["ndof"] // 6, [b])
        ms_f = timeit(lambda: plan.update_fint(fint))
        print(json.dumps({"config": "eval %s, %d elements" % (kind, b.ne), "matrices": [m for m in ("KC0", "KG", "M") if b.sizes[m]],
                          "ms": ms, "elements_per_s": b.ne / ms * 1e3, "store_GBps": b.ne * bytes_el / ms / 1e6,
                          "frac_of_6538.9": b.ne * bytes_el / ms / 1e6 / 6538.9, "update_fint_ms": ms_f}))
        del b, coo, plan


def run_indices(side):
    """First-call cost: the int64 row/col index arrays (the unrolled KC0r[k] = ...; KC0c[k] = ... blocks)."""
    case = meshes.plate_quad4(side, side)
    b = util.batch_from_case(case)
    for m in ("KC0", "KG", "M"):
        coo = b.fill_indices(m)
        ms = timeit(lambda m=m, coo=coo: b.fill_indices(m, coo=coo), steps=3, warm=1)
        n = b.ne * b.sizes[m]
        print(json.dumps({"config": "fill_indices %s %dx%d Quad4" % (m, side, side), "entries": n, "ms": ms,
                          "GBps": n * 16 / ms / 1e6, "frac_of_6538.9": n * 16 / ms / 1e6 / 6538.9}))
        del coo


def run_fint(side):
    """update_fint of every element of the north-star mesh, element kernel + plan gather."""
    case = meshes.plate_quad4(side, side)
    for kind in ("quad4", "quad4r"):
        case["kind"] = kind
        b = util.batch_from_case(case)
        nn = case["ndof"] // 6
        plan = AssemblyPlan("KC0", nn, [b])
        fint = torch.zeros(6 * nn, dtype=torch.float64, device=b.device)
        ms = timeit(lambda: plan.update_fint(fint))
        ms_sort = timeit(lambda: b.update_fint(fint), steps=3, warm=1)
        ne = case["conn"].shape[0]
        alg = ne * (32 + 24 + 48 + 2 * 192) + nn * 48 * 2
        print(json.dumps({"config": "update_fint %s %dx%d" % (kind, side, side), "elements": ne, "ms_plan_gather": ms,
                          "ms_sort_gather": ms_sort, "elements_per_s": ne / ms * 1e3, "algorithmic_GBps": alg / ms / 1e6,
                          "frac_of_6538.9": alg / ms / 1e6 / 6538.9}))
        del plan, b


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
    if "--mixed" in sys.argv:
        run_mixed(140 if "--small" in sys.argv else 1403)
        sys.exit(0)
    if "--kinds" in sys.argv:
        run_kinds(20000 if "--small" in sys.argv else 1000000)
        sys.exit(0)
    if "--indices" in sys.argv:
        run_indices(200 if "--small" in sys.argv else 2000)
        sys.exit(0)
    if "--fint" in sys.argv:
        run_fint(200 if "--small" in sys.argv else 2000)
        sys.exit(0)
    if "--aero" in sys.argv:
        run_aero(200 if "--small" in sys.argv else 2000)
        sys.exit(0)
    if "--dropin" in sys.argv:
        run_dropin(400 if "--small" in sys.argv else 2000)
        sys.exit(0)
    if "--cg" in sys.argv:
        run_cg(100 if "--small" in sys.argv else 2000, host_side=50 if "--small" in sys.argv else 400)
        sys.exit(0)
    if "--spmv" in sys.argv:
        run_spmv(200 if "--small" in sys.argv else 2000)
        sys.exit(0)
    small = "--small" in sys.argv
    f = 0.1 if small else 1.0
    if "--config2" in sys.argv:
        run("config2 BeamC arc 100k, KC0+M", meshes.arc_beamc(int(100001 * f)), ("KC0", "M0"), False)
        sys.exit(0)
    if "--config3" in sys.argv:
        run("config3 Quad4R cylinder 1M, KC0+KG_given_stress", meshes.cylinder_quad4r(int(1760 * f ** 0.5), int(571 * f ** 0.5)),
            ("KC0", "KGs"), True)
        sys.exit(0)
    if "--subsets" in sys.argv:   # one- and two-matrix calls on the north-star mesh (the calls that take several node pairs per CTA)
        side = int(2000 * f ** 0.5)
        for kind, mk in (("quad4", meshes.plate_quad4), ("quad4r", meshes.plate_quad4)):
            case = mk(side, side, kind=kind)
            run("%s %dx%d KC0" % (kind, side, side), case, ("KC0",), True)
            run("%s %dx%d KC0+KG" % (kind, side, side), case, ("KC0", "KG"), True)
            run("%s %dx%d KC0+M0" % (kind, side, side), case, ("KC0", "M0"), True)
        sys.exit(0)
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
