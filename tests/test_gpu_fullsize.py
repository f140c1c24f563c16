"""GPU, BASELINE.json full size (4.0 M Quad4): size-independent properties of the evaluated and assembled
matrices, plus the empty / single-element / ragged edge cases.  The oracle cannot run at this size in
seconds, so parity here is through invariants of the domain:
  * checksum conservation: sum(COO values) == sum(CSR values) for KC0, KG, M (assembly only adds)
  * symmetry: x.(K y) == y.(K x) through the device SpMV
  * self-consistency: K u == fint(u)  (Quad4: quad4.pyx:1174-1201 computes finte = KC0ve.ue)
  * rigid-body translation: K t == 0;  total mass: t.(M t) == intrho * area
  * linearity of KG in u;  fused path == two-pass path
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIDE = 2000


@pytest.fixture(scope="module")
def big():
    import gc
    import torch
    from pyfe3d_b200 import meshes
    from pyfe3d_b200.batch import AssemblyPlan, ElementBatch
    gc.collect()
    torch.cuda.empty_cache()     # the C ABI's own cudaMalloc calls cannot use blocks torch still caches
    case = meshes.plate_quad4(SIDE, SIDE)
    b = ElementBatch("quad4", case["conn"], case["x"], case["props"], u=case["u"])
    nn = case["ndof"] // 6
    plans = {m: AssemblyPlan(m, nn, [b]) for m in ("KC0", "KG", "M")}
    coo, csr = plans["KC0"].evaluate_assemble(KC0=True, KG=True, M=True)
    torch.cuda.synchronize()
    return dict(case=case, b=b, nn=nn, plans=plans, coo=coo, csr=csr)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def test_full_size_shapes(big):
    ne = SIDE * SIDE
    assert big["b"].ne == ne == 4_000_000
    assert big["coo"]["KC0"].v.numel() == 576 * ne          # > 2^31: beyond the reference's int32 init_k
    assert big["plans"]["KC0"].nnz == big["csr"]["KC0"].numel()
    # structured plate: interior nodes couple to 9 nodes, edges 6, corners 4
    n = SIDE + 1
    nblk = 9 * (n - 2) ** 2 + 6 * 4 * (n - 2) + 4 * 4
    assert big["plans"]["KC0"].nnz == 36 * nblk and big["csr"]["KG"].numel() == 9 * nblk
    assert big["csr"]["M"].numel() == 30 * nblk


@pytest.mark.parametrize("m", ["KC0", "KG", "M"])
def test_checksum_conservation(big, m):
    s_coo = big["coo"][m].v.sum(dtype=__import__("torch").float64)
    s_csr = big["csr"][m].sum()
    scale = big["coo"][m].v.abs().sum()
    assert abs(float(s_coo - s_csr)) <= 1e-11 * float(scale)


def test_symmetry_and_fint_and_rigid_body(big):
    import torch
    from pyfe3d_b200.batch import spmv
    case, b, nn = big["case"], big["b"], big["nn"]
    indptr, indices = big["plans"]["KC0"].pattern()
    K = big["csr"]["KC0"]
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(6 * nn, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(6 * nn, dtype=torch.float64, device="cuda", generator=g)
    Kx, Ky = spmv(indptr, indices, K, x), spmv(indptr, indices, K, y)
    assert abs(float(torch.dot(y, Kx) - torch.dot(x, Ky))) <= 1e-10 * float(torch.dot(x, Kx).abs())
    # K u == fint(u)
    fint = torch.zeros(6 * nn, dtype=torch.float64, device="cuda")
    b.update_fint(fint)
    Ku = spmv(indptr, indices, K, b.u)
    assert _rel(Ku, fint) <= 1e-9
    # rigid translation produces no force
    t = torch.zeros(6 * nn, dtype=torch.float64, device="cuda")
    t[0::6], t[1::6], t[2::6] = 0.3, -0.2, 0.7
    Kt = spmv(indptr, indices, K, t)
    assert float(Kt.abs().max()) <= 1e-9 * float(Kx.abs().max())
    # total mass of the consistent mass matrix = intrho * plate area (unit plate)
    mp, mi = big["plans"]["M"].pattern()
    Mt = spmv(mp, mi, big["csr"]["M"], t)
    rho0 = float(case["props"][0, 24])
    want = rho0 * 1.0 * (0.3 ** 2 + 0.2 ** 2 + 0.7 ** 2)
    assert abs(float(torch.dot(t, Mt)) - want) <= 1e-10 * want


def test_kg_linear_in_u_and_fused_equals_twopass(big):
    import torch
    b, plans = big["b"], big["plans"]
    _, csr2 = plans["KC0"].evaluate_assemble(KG=True, u=2.0 * b.u, write_coo=False)
    assert _rel(csr2["KG"], 2.0 * big["csr"]["KG"]) <= 1e-12
    del csr2
    two = b.evaluate(KG=True, indices=False)
    ref = plans["KG"].assemble(two["KG"].v)
    assert _rel(big["csr"]["KG"], ref) <= 1e-12
    assert _rel(big["coo"]["KG"].v, two["KG"].v) <= 1e-12


def test_edge_cases_empty_single_and_ragged():
    import torch
    from pyfe3d_b200 import meshes
    from pyfe3d_b200.batch import AssemblyPlan, ElementBatch
    case = meshes.plate_quad4(1, 1)
    # empty batch: every call is a no-op
    b0 = ElementBatch("quad4", np.zeros((0, 4), np.int64), case["x"], case["props"], u=case["u"])
    out = b0.evaluate(KC0=True, KG=True, M=True)
    assert all(v.v.numel() == 0 for v in out.values())
    # single element: one 24x24 dense block, 16 node-pair blocks
    b1 = ElementBatch("quad4", case["conn"], case["x"], case["props"], u=case["u"])
    p1 = AssemblyPlan("KC0", 4, [b1])
    coo, csr = p1.evaluate_assemble(KC0=True)
    assert p1.nnz == 576
    A = p1.to_scipy(csr["KC0"]).toarray()
    assert np.abs(A - A.T).max() <= 1e-12 * np.abs(A).max()
    assert abs(float(coo["KC0"].v.sum()) - A.sum()) <= 1e-11 * np.abs(A).sum()
    # ragged valence: an L-shaped patch (nodes with 1, 2, 3 and 4 incident elements) incl. unused nodes
    c = meshes.plate_quad4(3, 3)
    keep = np.array([0, 1, 2, 3, 4, 6])        # drop three elements -> ragged, one node without elements
    br = ElementBatch("quad4", c["conn"][keep], c["x"], c["props"], u=c["u"])
    pr = AssemblyPlan("KC0", c["ndof"] // 6, [br])
    coo, csr = pr.evaluate_assemble(KC0=True)
    two = br.update_KC0(update_KC0v_only=1)
    ref = pr.assemble(two.v)
    assert float((csr["KC0"] - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    indptr, _ = pr.pattern()
    rows = (indptr[1:] - indptr[:-1]).cpu().numpy().reshape(-1, 6)
    assert (rows == 0).all(axis=1).sum() >= 1       # the orphan node has empty rows


def test_plan_spmv_and_aero_at_full_size(big):
    """Block SpMV against the per-entry CSR SpMV on the 1.3 G-nonzero matrix; K u == fint through it; the
    piston-theory matrices at 4 M elements: CA == -KA_gamma, KA_gamma == consistent-mass H (t.KA_gamma t = area n_z^2)."""
    import torch
    from pyfe3d_b200.batch import AssemblyPlan, spmv
    b, nn, plans, K = big["b"], big["nn"], big["plans"], big["csr"]["KC0"]
    indptr, indices = plans["KC0"].pattern()
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(6 * nn, dtype=torch.float64, device="cuda", generator=g)
    y_blk = plans["KC0"].spmv(K, x)
    y_csr = spmv(indptr, indices, K, x)
    assert _rel(y_blk, y_csr) <= 1e-13
    fint = torch.zeros(6 * nn, dtype=torch.float64, device="cuda")
    plans["KC0"].update_fint(fint)
    assert _rel(plans["KC0"].spmv(K, b.u), fint) <= 1e-9
    free = (torch.rand(6 * nn, device="cuda", generator=g) > 0.1).to(torch.uint8)
    ym = plans["KC0"].spmv(K, x, free=free)
    f64 = free.to(torch.float64)
    assert _rel(ym, spmv(indptr, indices, K, x * f64) * f64) <= 1e-13
    del y_blk, y_csr, ym, fint
    aero = b.evaluate_aero(KA_beta=True, KA_gamma=True, CA=True, indices=False)
    assert torch.equal(aero["CA"].v, -aero["KA_gamma"].v)
    pg = AssemblyPlan("KA_gamma", nn, [b])
    G = pg.assemble(aero["KA_gamma"].v)
    t = torch.zeros(6 * nn, dtype=torch.float64, device="cuda")
    t[0::6], t[1::6], t[2::6] = 0.3, -0.2, 0.7
    # sum_ab H_ab = area; global block = H_ab z z^T  =>  t.G t = area * (t.z)^2 with z the plate normal
    from pyfe3d_b200 import meshes
    z = meshes.fixed_rotation(0)[:, 2]
    want = 1.0 * float(np.dot([0.3, -0.2, 0.7], z)) ** 2
    assert abs(float(torch.dot(t, pg.spmv(G, t))) - want) <= 1e-10 * abs(want)
    # KA_beta: rows sum to zero over b (sum_b N_b,x = 0): a rigid translation sees no aerodynamic stiffness
    pb = AssemblyPlan("KA_beta", nn, [b])
    Bm = pb.assemble(aero["KA_beta"].v)
    assert float(pb.spmv(Bm, t).abs().max()) <= 1e-9 * float(Bm.abs().max())


def test_config4_tria3r_4M_fused_properties():
    """Config 4 at full size (4.0 M Tria3R, distorted plate, KC0 + M mtype 1) through the fused triangle kernel:
    checksum conservation, symmetry, rigid-body null space, total mass, fused == two-pass."""
    import torch
    from pyfe3d_b200 import meshes
    from pyfe3d_b200.batch import AssemblyPlan
    from tests import util
    case = meshes.plate_tria3r(1415, 1415)
    b = util.batch_from_case(case)
    ne, nn = case["conn"].shape[0], case["ndof"] // 6
    assert ne == 2 * 1415 * 1415
    plan = AssemblyPlan("KC0", nn, [b])
    coo, csr = plan.evaluate_assemble(KC0=True, M=True, mtype=1)
    for m in ("KC0", "M"):
        s_coo, s_csr, scale = coo[m].v.sum(), csr[m].sum(), coo[m].v.abs().sum()
        assert abs(float(s_coo - s_csr)) <= 1e-11 * float(scale)
    K = csr["KC0"]
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(6 * nn, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(6 * nn, dtype=torch.float64, device="cuda", generator=g)
    Kx, Ky = plan.spmv(K, x), plan.spmv(K, y)
    assert abs(float(torch.dot(y, Kx) - torch.dot(x, Ky))) <= 1e-10 * float(torch.dot(x, Kx).abs())
    t = torch.zeros(6 * nn, dtype=torch.float64, device="cuda")
    t[0::6], t[1::6], t[2::6] = 0.3, -0.2, 0.7
    assert float(plan.spmv(K, t).abs().max()) <= 1e-9 * float(Kx.abs().max())
    pm = plan._sibling("M", 1)
    rho0 = float(case["props"][0, 24])
    want = rho0 * 0.3 * 0.5 * (0.3 ** 2 + 0.2 ** 2 + 0.7 ** 2)      # plate a x b = 0.3 x 0.5
    assert abs(float(torch.dot(t, pm.spmv(csr["M"], t))) - want) <= 1e-9 * want
    del Kx, Ky, x, y
    two = b.evaluate(KC0=True, M=True, mtype=1, indices=False)
    for m in ("KC0", "M"):
        assert _rel(coo[m].v, two[m].v) <= 1e-12
        ref = plan._sibling(m, 1).assemble(two[m].v)
        assert _rel(csr[m], ref) <= 1e-12
        del ref


def test_config3_quad4r_1M_fused_properties():
    """Config 3 at full size (1.0 M Quad4R cylinder, material axes, KC0 + KG_given_stress): checksum, symmetry,
    fused == two-pass."""
    import torch
    from pyfe3d_b200 import meshes
    from pyfe3d_b200.batch import AssemblyPlan
    from tests import util
    case = meshes.cylinder_quad4r(1760, 571)
    b = util.batch_from_case(case)
    nn = case["ndof"] // 6
    st = case.get("stress", (0., 0., 1.))
    plan = AssemblyPlan("KC0", nn, [b])
    coo, csr = plan.evaluate_assemble(KC0=True, KG_given_stress=st)
    for m in ("KC0", "KG"):
        assert abs(float(coo[m].v.sum() - csr[m].sum())) <= 1e-11 * float(coo[m].v.abs().sum())
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(6 * nn, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(6 * nn, dtype=torch.float64, device="cuda", generator=g)
    for m in ("KC0", "KG"):
        p = plan._sibling(m, 0)
        Ax, Ay = p.spmv(csr[m], x), p.spmv(csr[m], y)
        assert abs(float(torch.dot(y, Ax) - torch.dot(x, Ay))) <= 1e-10 * float(torch.dot(x, Ax).abs() + Ax.abs().max())
    two = b.evaluate(KC0=True, KG_given_stress=st, indices=False)
    for m in ("KC0", "KG"):
        assert _rel(coo[m].v, two[m].v) <= 1e-12
        assert _rel(csr[m], plan._sibling(m, 0).assemble(two[m].v)) <= 1e-12
