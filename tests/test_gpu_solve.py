"""GPU: the native device solve (pf3_plan_cg), the scaled operator (pf3_plan_spmv_scaled) and the explicit
boundary-condition partition K[bu,:][:,bu] (pf3_csr_compact_*) -- SURVEY 8(f) ranks 1-2 -- against scipy on the same
matrices, in the forms the reference scripts use them (tests/test_quad4_static_point_load.py:84-104,
tests/test_quad4r_linear_buckling_plate.py:135-146,172-180)."""
import numpy as np
import pytest

from tests import cases, util

pytestmark = pytest.mark.gpu


def _plate(kind="quad4", nx=11, ny=9, seed=3):
    """A connected distorted shell mesh, its KC0 / M plans and values, and a clamped-edge dof mask."""
    import torch  # noqa: F401
    from pyfe3d_b200.batch import AssemblyPlan
    case = cases.shell_mesh(kind, nx, ny, seed=seed)
    b = util.batch_from_case(case)
    n = case["ndof"]
    pk = AssemblyPlan("KC0", n // 6, [b])
    pm = AssemblyPlan("M", n // 6, [b])
    K = pk.assemble(b.update_KC0(update_KC0v_only=1).v)
    M = pm.assemble(b.update_M(mtype=0, indices=False).v)
    conn = np.asarray(case["conn"])
    # constrain every dof of the nodes of the first 2*ny elements' first nodes (a clamped strip), and drilling nowhere
    fixed_nodes = np.unique(conn[: 2 * ny])
    free = np.ones(n, np.uint8)
    for d in range(6):
        free[6 * fixed_nodes + d] = 0
    return case, pk, K, pm, M, free


def test_native_cg_matches_spsolve():
    import torch
    from scipy.sparse.linalg import spsolve
    from pyfe3d_b200.solve import plan_cg_native
    case, pk, K, pm, M, free = _plate()
    n = case["ndof"]
    rng = np.random.default_rng(1)
    f = rng.normal(size=n)
    bu = free.astype(bool)
    A = pk.to_scipy(K).tocsc()
    want = np.zeros(n)
    want[bu] = spsolve(A[bu, :][:, bu], f[bu])
    # (Jacobi-CG on a thin coupled laminate needs many times ndof iterations: give it room)
    x, it, status, res, bn = plan_cg_native(pk, K, torch.as_tensor(f).cuda(), free=torch.as_tensor(free).cuda(),
                                            rtol=1e-12, check_every=8, maxiter=500000)
    assert status == 0 and it > 0
    assert res <= 1e-12 * bn
    got = x.cpu().numpy()
    assert np.all(got[~bu] == 0.)
    assert np.abs(got - want).max() <= 1e-7 * np.abs(want).max()
    # bit-reproducible: fixed-order reductions
    x2, it2, *_ = plan_cg_native(pk, K, torch.as_tensor(f).cuda(), free=torch.as_tensor(free).cuda(), rtol=1e-12,
                                 check_every=8, maxiter=500000)
    assert it2 == it and torch.equal(x, x2)
    # warm start from the solution converges at once
    x3, it3, st3, *_ = plan_cg_native(pk, K, torch.as_tensor(f).cuda(), free=torch.as_tensor(free).cuda(), rtol=1e-7,
                                      x0=x, maxiter=500000)
    assert st3 == 0 and it3 <= 2
    assert np.abs(x3.cpu().numpy() - want).max() <= 1e-7 * np.abs(want).max()


def test_native_cg_is_the_reference_scaled_cg():
    """D = diag(Kuu)^-1/2; cg(D Kuu D, D fu, atol) of tests/test_quad4r_linear_buckling_plate.py:135-146: same stopping
    test (scaled norm, atol), same solution."""
    import scipy.sparse as sp
    import torch
    from scipy.sparse.linalg import cg
    from pyfe3d_b200.solve import plan_cg_native
    case, pk, K, pm, M, free = _plate("quad4r", 17, 13, seed=5)
    n = case["ndof"]
    bu = free.astype(bool)
    f = np.zeros(n)
    f[2::6] = 1.0
    Kuu = pk.to_scipy(K).tocsc()[bu, :][:, bu]
    dis = 1.0 / np.sqrt(np.maximum(Kuu.diagonal(), 1e-30))
    D = sp.diags(dis)
    fs = D @ f[bu]
    atol = 1e-9 * np.linalg.norm(fs)
    us, out = cg(D @ Kuu @ D, fs, atol=atol, rtol=0., maxiter=20000)
    assert out == 0
    want = np.zeros(n)
    want[bu] = D @ us
    x, it, status, res, bn = plan_cg_native(pk, K, torch.as_tensor(f).cuda(), free=torch.as_tensor(free).cuda(),
                                            rtol=0., atol=atol, scaled_norm=True)
    assert status == 0
    assert res <= atol and abs(bn - np.linalg.norm(fs)) <= 1e-12 * bn
    assert np.abs(x.cpu().numpy() - want).max() <= 1e-6 * np.abs(want).max()


def test_native_cg_sum_of_operators_and_failure_modes():
    import torch
    from scipy.sparse.linalg import spsolve
    from pyfe3d_b200.solve import plan_cg_native
    case, pk, K, pm, M, free = _plate()
    n = case["ndof"]
    bu = free.astype(bool)
    f = np.random.default_rng(2).normal(size=n)
    sigma = -3.0e3
    A = (pk.to_scipy(K) - sigma * pm.to_scipy(M)).tocsc()
    want = np.zeros(n)
    want[bu] = spsolve(A[bu, :][:, bu], f[bu])
    ft, fr = torch.as_tensor(f).cuda(), torch.as_tensor(free).cuda()
    x, it, status, *_ = plan_cg_native(pk, K, ft, free=fr, rtol=1e-12, extra=[(pm, M, -sigma)], maxiter=500000)
    assert status == 0
    assert np.abs(x.cpu().numpy() - want).max() <= 1e-7 * np.abs(want).max()
    # maxiter reached is reported, not an exception; x is finite
    x, it, status, *_ = plan_cg_native(pk, K, ft, free=fr, rtol=1e-14, maxiter=3)
    assert status == 1 and it == 3 and bool(torch.isfinite(x).all())
    # an indefinite operator (-K) breaks down cleanly
    x, it, status, *_ = plan_cg_native(pk, -K, ft, free=fr, rtol=1e-14)
    assert status == 2 and bool(torch.isfinite(x).all())
    # zero right-hand side: converged at iteration 0
    x, it, status, *_ = plan_cg_native(pk, K, torch.zeros_like(ft), free=fr)
    assert status == 0 and it == 0 and float(x.abs().max()) == 0.


def test_scaled_operator_matches_scipy():
    import scipy.sparse as sp
    import torch
    case, pk, K, pm, M, free = _plate()
    n = case["ndof"]
    A = pk.to_scipy(K)
    d = pk.diagonal(K).cpu().numpy()
    s = 1.0 / np.sqrt(np.maximum(d, 1e-30))
    x = np.random.default_rng(4).normal(size=n)
    P = sp.diags(free.astype(float))
    S = sp.diags(s)
    ref = S @ (P @ (A @ (P @ (S @ x))))
    y = pk.spmv_scaled(K, torch.as_tensor(s).cuda(), torch.as_tensor(x).cuda(), free=torch.as_tensor(free).cuda())
    assert np.abs(y.cpu().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("matrix", ["KC0", "M"])
@pytest.mark.parametrize("shard", [False, True])
def test_compact_csr_is_scipy_extraction(matrix, shard):
    """K[bu,:][:,bu] (tests/test_quad4_static_point_load.py:84-99): pattern bit-exact, values bit-exact (a pure copy)."""
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    from pyfe3d_b200.solve import plan_compact
    case = cases.shell_mesh("quad4", 14, 11, seed=9)
    b = util.batch_from_case(case)
    n = case["ndof"]
    nn = n // 6
    rng_ = (nn // 3, nn - 7) if shard else None
    plan = AssemblyPlan(matrix, nn, [b], node_range=rng_)
    coo = b.update_KC0(update_KC0v_only=1) if matrix == "KC0" else b.update_M(indices=False)
    vals = plan.assemble(coo.v)
    rng = np.random.default_rng(12)
    free = (rng.random(n) > 0.3).astype(np.uint8)
    free[2::6] = 1
    bu = free.astype(bool)
    lo = 6 * plan.node_begin
    A = plan.to_scipy(vals).tocsr()
    want = A[bu[lo:lo + plan.nrows], :][:, bu].tocsr()
    want.sort_indices()
    (ip, ix, v), pat = plan_compact(plan, vals, torch.as_tensor(free).cuda())
    assert np.array_equal(ip.cpu().numpy(), want.indptr)
    assert np.array_equal(ix.cpu().numpy(), want.indices)
    assert np.array_equal(v.cpu().numpy(), want.data)
    # values-only refresh through the cached pattern
    (ip2, ix2, v2), _ = plan_compact(plan, 2.0 * vals, torch.as_tensor(free).cuda(), pattern=pat)
    assert ip2 is ip and ix2 is ix
    assert np.array_equal(v2.cpu().numpy(), 2.0 * want.data)


def test_compact_csr_edge_masks():
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    from pyfe3d_b200.solve import plan_compact
    case = cases.shell_mesh("quad4", 5, 4, seed=1)
    b = util.batch_from_case(case)
    n = case["ndof"]
    plan = AssemblyPlan("KC0", n // 6, [b])
    vals = plan.assemble(b.update_KC0(update_KC0v_only=1).v)
    A = plan.to_scipy(vals).tocsr()
    A.sort_indices()
    # everything free: the matrix itself
    (ip, ix, v), _ = plan_compact(plan, vals, torch.ones(n, dtype=torch.uint8).cuda())
    assert np.array_equal(ip.cpu().numpy(), A.indptr) and np.array_equal(ix.cpu().numpy(), A.indices)
    assert np.array_equal(v.cpu().numpy(), A.data)
    # nothing free: an empty 0 x 0 matrix
    (ip, ix, v), _ = plan_compact(plan, vals, torch.zeros(n, dtype=torch.uint8).cuda())
    assert ip.numel() == 1 and int(ip[0]) == 0 and ix.numel() == 0 and v.numel() == 0


@pytest.mark.parametrize("masked", [False, True])
def test_compact_upper_triangle_is_scipy_triu(masked):
    """PF3_COMPACT_UPPER: one triangle of the (symmetric) matrix, optionally of K[bu,:][:,bu]."""
    import scipy.sparse as sp
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    from pyfe3d_b200.solve import plan_compact
    case = cases.shell_mesh("quad4", 9, 7, seed=6)
    b = util.batch_from_case(case)
    n = case["ndof"]
    plan = AssemblyPlan("KC0", n // 6, [b])
    vals = plan.assemble(b.update_KC0(update_KC0v_only=1).v)
    A = plan.to_scipy(vals).tocsr()
    free = None
    if masked:
        free = (np.random.default_rng(3).random(n) > 0.25).astype(np.uint8)
        bu = free.astype(bool)
        A = A[bu, :][:, bu]
    want = sp.triu(A, format="csr")
    want.sort_indices()
    (ip, ix, v), pat = plan_compact(plan, vals, None if free is None else torch.as_tensor(free).cuda(), upper=True)
    assert np.array_equal(ip.cpu().numpy(), want.indptr)
    assert np.array_equal(ix.cpu().numpy(), want.indices)
    assert np.array_equal(v.cpu().numpy(), want.data)
    assert v.numel() < 0.6 * vals.numel()
    # the triangle determines the matrix: K = U + U^T - diag(U) to rounding (symmetry of the assembled values)
    U = sp.csr_matrix((v.cpu().numpy(), ix.cpu().numpy(), ip.cpu().numpy()), shape=want.shape)
    full = U + U.T - sp.diags(U.diagonal())
    assert abs(full - A).max() <= 1e-12 * abs(A).max()
    # values-only refresh without index arrays
    (ip3, ix3, v3), _ = plan_compact(plan, vals, None if free is None else torch.as_tensor(free).cuda(), upper=True,
                                     want_indices=False)
    assert ix3 is None and np.array_equal(v3.cpu().numpy(), want.data)


def test_cg_shard_step_kernels_match_torch():
    """pf3_cg_shard_dot / _update / _dir (the fused vector kernels of the row-sharded CG) against the same algebra in torch."""
    import torch
    from pyfe3d_b200 import _cabi
    from pyfe3d_b200.batch import context
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(5)
    for n in (1, 37, 100003, 3000001):
        p, ap, r, x = (torch.randn(n, dtype=torch.float64, generator=g).to(dev) for _ in range(4))
        minv = torch.rand(n, dtype=torch.float64, generator=g).to(dev) + 0.5
        sc = torch.tensor([0., 0., 0., 1.7], dtype=torch.float64, device=dev)
        work = torch.zeros(_cabi.cg_shard_work_bytes() // 8 + 1, dtype=torch.float64, device=dev)
        ctx = context(dev)
        p0, x0, r0 = p.clone(), x.clone(), r.clone()
        for _ in range(2):   # twice: the tickets must have been reset
            ctx.cg_shard_dot(n, p.data_ptr(), ap.data_ptr(), sc.data_ptr(), work.data_ptr())
        want_pap = torch.dot(p0, ap)
        assert abs(float(sc[0]) - float(want_pap)) <= 1e-12 * float(p0.abs() @ ap.abs())
        sc[0] = 2.5                                                     # as if all-reduced
        ctx.cg_shard_update(n, p.data_ptr(), ap.data_ptr(), minv.data_ptr(), x.data_ptr(), r.data_ptr(), sc.data_ptr(),
                            work.data_ptr())
        alpha = 1.7 / 2.5
        xr, rr = x0 + alpha * p0, r0 - alpha * ap
        assert torch.allclose(x, xr, rtol=1e-14, atol=1e-14) and torch.allclose(r, rr, rtol=1e-14, atol=1e-14)
        assert abs(float(sc[1]) - float((rr * minv) @ rr)) <= 1e-12 * float((rr * minv) @ rr)
        assert abs(float(sc[2]) - float(rr @ rr)) <= 1e-12 * float(rr @ rr)
        rzn = float(sc[1])
        ctx.cg_shard_dir(n, r.data_ptr(), minv.data_ptr(), p.data_ptr(), sc.data_ptr(), work.data_ptr())
        assert torch.allclose(p, minv * rr + (rzn / 1.7) * p0, rtol=1e-13, atol=1e-13)
        assert float(sc[3]) == rzn                                      # r.z rolled over for the next iteration


def test_cg_shard_step_kernels_freeze_on_zero_denominators():
    """p.Ap = 0 / r.z = 0 (an exactly converged system iterated on inside a batch) must not spread NaN."""
    import torch
    from pyfe3d_b200 import _cabi
    from pyfe3d_b200.batch import context
    dev = torch.device("cuda", 0)
    n = 1000
    p, ap, minv = torch.ones(n, dtype=torch.float64, device=dev), torch.zeros(n, dtype=torch.float64, device=dev), \
        torch.ones(n, dtype=torch.float64, device=dev)
    x, r = torch.full((n,), 3.0, dtype=torch.float64, device=dev), torch.zeros(n, dtype=torch.float64, device=dev)
    sc = torch.zeros(4, dtype=torch.float64, device=dev)
    work = torch.zeros(_cabi.cg_shard_work_bytes() // 8 + 1, dtype=torch.float64, device=dev)
    ctx = context(dev)
    ctx.cg_shard_dot(n, p.data_ptr(), ap.data_ptr(), sc.data_ptr(), work.data_ptr())
    ctx.cg_shard_update(n, p.data_ptr(), ap.data_ptr(), minv.data_ptr(), x.data_ptr(), r.data_ptr(), sc.data_ptr(),
                        work.data_ptr())
    ctx.cg_shard_dir(n, r.data_ptr(), minv.data_ptr(), p.data_ptr(), sc.data_ptr(), work.data_ptr())
    assert torch.isfinite(x).all() and torch.isfinite(p).all() and torch.isfinite(sc).all()
    assert torch.equal(x, torch.full_like(x, 3.0)) and float(sc[2]) == 0.0
