"""GPU: block-structured SpMV / diagonal on the plan's own CSR layout (pf3_plan_spmv, pf3_plan_diagonal) against
scipy on the same matrix, with and without the boundary-condition mask and on a row shard; plan-based Jacobi-CG on
config 1."""
import json
import os

import numpy as np
import pytest

from tests import cases, configs, util

pytestmark = pytest.mark.gpu


def _plan_and_values(case, matrix, mtype=0, node_range=None):
    from pyfe3d_b200.batch import AssemblyPlan
    b = util.batch_from_case(case)
    nn = case["ndof"] // 6
    plan = AssemblyPlan(matrix, nn, [b], node_range=node_range, mtype=mtype)
    if matrix == "KC0":
        coo = b.update_KC0(update_KC0v_only=1)
    elif matrix == "KG":
        coo = b.update_KG(update_KGv_only=1)
    else:
        coo = b.update_M(mtype=mtype, indices=False)
    return plan, plan.assemble(coo.v)


@pytest.mark.parametrize("matrix,mtype", [("KC0", 0), ("KG", 0), ("M", 0), ("M", 1), ("M", 2)])
@pytest.mark.parametrize("shard", [False, True])
def test_plan_spmv_matches_scipy(matrix, mtype, shard):
    import torch
    case = cases.shell_mesh("quad4", 23, 17, seed=11)
    n = case["ndof"]
    nn = n // 6
    rng_ = (nn // 4, nn - 5) if shard else None
    plan, vals = _plan_and_values(case, matrix, mtype, rng_)
    A = plan.to_scipy(vals)
    rng = np.random.default_rng(3)
    x = rng.normal(size=n)
    y = plan.spmv(vals, torch.as_tensor(x).cuda()).cpu().numpy()
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()
    free = (rng.random(n) > 0.2).astype(np.uint8)
    ym = plan.spmv(vals, torch.as_tensor(x).cuda(), free=torch.as_tensor(free).cuda()).cpu().numpy()
    lo = 6 * plan.node_begin
    refm = (A @ (x * free)) * free[lo:lo + plan.nrows]
    assert np.abs(ym - refm).max() <= 1e-13 * np.abs(ref).max()
    d = plan.diagonal(vals).cpu().numpy()
    dref = np.asarray(A[:, lo:lo + plan.nrows].diagonal())
    assert np.array_equal(d, dref)


@pytest.mark.parametrize("kind", ["tria3r", "beamc", "truss"])
def test_plan_spmv_other_kinds(kind):
    import torch
    case = cases.shell_mesh("tria3r", 9, 8, seed=2) if kind == "tria3r" else cases.line_chain(kind, 40, seed=4)
    plan, vals = _plan_and_values(case, "KC0")
    A = plan.to_scipy(vals)
    x = np.random.default_rng(5).normal(size=case["ndof"])
    y = plan.spmv(vals, torch.as_tensor(x).cuda()).cpu().numpy()
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.array_equal(plan.diagonal(vals).cpu().numpy(), A.diagonal())


def test_plan_spmv_mixed_groups():
    """Quad4 skin + BeamC stiffeners in one plan (different masks -> the general row loop)."""
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    cs = configs.build("stiffened_panel")
    bs = [util.batch_from_case(c) for c in cs]
    nn = cs[0]["ndof"] // 6
    for matrix in ("KC0", "M"):
        plan = AssemblyPlan(matrix, nn, bs)
        vs = [b.update_KC0(update_KC0v_only=1).v if matrix == "KC0" else b.update_M(indices=False).v for b in bs]
        vals = plan.assemble(torch.cat(vs))
        A = plan.to_scipy(vals)
        x = np.random.default_rng(8).normal(size=6 * nn)
        y = plan.spmv(vals, torch.as_tensor(x).cuda()).cpu().numpy()
        ref = A @ x
        assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()


def test_plan_cg_solves_static_config():
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    from pyfe3d_b200.solve import plan_cg_solve
    gold = json.load(open(os.path.join(util.GOLDEN_DIR, "config_scalars.json")))
    c = configs.build("quad4_static")[0]
    b = util.batch_from_case(c)
    n = c["ndof"]
    plan = AssemblyPlan("KC0", n // 6, [b])
    _, csr = plan.evaluate_assemble(KC0=True, write_coo=False)
    m = c["meta"]
    X = c["x"].reshape(-1, 3)
    x, y = X[:, 0], X[:, 1]
    edge = np.isclose(x, 0.) | np.isclose(x, m["a"]) | np.isclose(y, 0.) | np.isclose(y, m["b"])
    bk = np.zeros(n, bool)
    bk[2::6] = edge
    bk[0::6] = True
    bk[1::6] = True
    bk[5::6] = True
    f = np.zeros(n)
    f[2::6][np.isclose(x, m["a"] / 2) & np.isclose(y, m["b"] / 2)] = 1.
    u, info = plan_cg_solve(plan, csr["KC0"], torch.as_tensor(f).cuda(),
                            free=torch.as_tensor((~bk).astype(np.uint8)).cuda(), rtol=1e-13)
    assert info > 0
    want = gold["quad4_static"]["w_max"]
    assert abs(float(u[2::6].max()) - want) <= 1e-8 * abs(want)


def test_shift_invert_operator_reproduces_beam_frequencies():
    """Config 2 (BeamC curved cantilever, tests/test_beamc_natural_freq_curved.py): eigsh in shift-invert mode with
    OPinv = device Jacobi-CG on K + M applied as two block SpMVs; first three natural frequencies against the
    reference-derived values at 1e-8."""
    import torch
    from scipy.sparse.linalg import eigsh
    from pyfe3d_b200.batch import AssemblyPlan
    from pyfe3d_b200.solve import shift_invert_operator
    gold = json.load(open(os.path.join(util.GOLDEN_DIR, "config_scalars.json")))["beamc_freq"]
    c = configs.build("beamc_freq")[0]
    b = util.batch_from_case(c)
    n = c["ndof"]
    pk = AssemblyPlan("KC0", n // 6, [b])
    pm = AssemblyPlan("M", n // 6, [b])
    K = pk.assemble(b.update_KC0(update_KC0v_only=1).v)
    M = pm.assemble(b.update_M(mtype=0, indices=False).v)
    bu = np.ones(n, bool)
    bu[:6] = False                       # clamped at the first node
    Kuu = pk.to_scipy(K).tocsc()[bu, :][:, bu]
    Muu = pm.to_scipy(M).tocsc()[bu, :][:, bu]
    op = shift_invert_operator(pk, K, pm, M, -1., bu.astype(np.uint8), rtol=1e-14)
    vals, _ = eigsh(A=Kuu, M=Muu, sigma=-1., which="LM", k=3, tol=0, OPinv=op)
    om = np.sqrt(np.sort(vals))
    for i, k in enumerate(("omega1", "omega2", "omega3")):
        assert abs(om[i] - gold[k]) <= 1e-8 * gold[k], (k, om[i], gold[k])
    assert op.stats["solves"] > 0


def test_spmv_csr_arbitrary_rows():
    """pf3_spmv_csr on a CSR matrix with odd / empty / single-entry rows (the paired 16-byte path starts at the first
    even position of a row and hands an odd first or last entry to single lanes)."""
    import scipy.sparse as sp
    import torch
    from pyfe3d_b200.batch import spmv
    rng = np.random.default_rng(21)
    n = 777
    A = sp.random(n, n, density=0.02, random_state=5, format="lil")
    A[5, :] = 0.
    A[6, :] = 0.
    A[7, :] = 0.
    A[7, 3] = 2.5                       # single entry
    for r in (11, 12, 13):
        A[r, rng.choice(n, size=65 + r, replace=False)] = rng.normal(size=65 + r)   # longer than one pass of the warp
    A = A.tocsr()
    A.sort_indices()
    x = rng.normal(size=n)
    y = spmv(torch.as_tensor(A.indptr.astype(np.int64)).cuda(), torch.as_tensor(A.indices.astype(np.int64)).cuda(),
             torch.as_tensor(A.data).cuda(), torch.as_tensor(x).cuda()).cpu().numpy()
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1.)
    assert y[5] == 0. and y[6] == 0. and y[7] == 2.5 * x[3]
