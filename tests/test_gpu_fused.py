"""GPU: the fused node-centric evaluate+assemble kernel against the reference goldens (COO values)
and scipy's sum-duplicates of the reference COO (CSR values)."""
import numpy as np
import pytest

from oracle import driver
from tests import cases, util

pytestmark = pytest.mark.gpu


def _check(case, ref, mtype=0, stress=None, node_range=None):
    import scipy.sparse as sp
    from pyfe3d_b200.batch import AssemblyPlan
    b = util.batch_from_case(case)
    n = case["ndof"]
    nn = n // 6
    plan = AssemblyPlan("KC0", nn, [b], node_range=node_range)
    kw = dict(KC0=True, M=True, mtype=mtype, indices=True)
    if stress is None:
        kw["KG"] = True
    else:
        kw["KG_given_stress"] = stress
    coo, csr = plan.evaluate_assemble(**kw)
    ne = case["conn"].shape[0]
    keys = {"KC0": "KC0", "KG": "KG" if stress is None else "KGs", "M": "M%d" % mtype}
    lo, hi = (0, nn) if node_range is None else node_range
    for name, rk in keys.items():
        r, c, v = ref[rk]
        if node_range is None:   # with a shard only the owned nodes' slabs are written
            assert np.array_equal(coo[name].r.cpu().numpy(), r) and np.array_equal(coo[name].c.cpu().numpy(), c)
            assert util.block_relerr(coo[name].v.cpu().numpy(), v, ne) <= util.TOL_VALUES, name
        p = AssemblyPlan(name, nn, [b], node_range=node_range, mtype=mtype)
        A = p.to_scipy(csr[name])
        S = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
        S.sum_duplicates()
        S.sort_indices()
        S = S[6 * lo:6 * hi]
        assert np.array_equal(A.indptr, S.indptr) and np.array_equal(A.indices, S.indices), name
        assert np.abs(A.data - S.data).max() <= util.TOL_CSR * np.abs(S.data).max(), name


@pytest.mark.parametrize("name", ["quad4_mesh", "quad4r_mesh", "quad4_soup", "quad4r_soup", "quad4_soup_thick",
                                  "tria3r_mesh", "tria3r_soup", "tria3r_soup_thick"])
@pytest.mark.parametrize("mtype", [0, 1, 2])
def test_fused_matches_reference_golden(name, mtype):
    case, ref = util.load_golden(name)
    _check(case, ref, mtype=mtype)


@pytest.mark.parametrize("name", ["quad4r_mesh", "tria3r_mesh"])
def test_fused_given_stress_and_shard(name):
    case, ref = util.load_golden(name)
    _check(case, ref, stress=case["stress"])
    nn = case["ndof"] // 6
    _check(case, ref, node_range=(nn // 3, nn - 2))


@pytest.mark.parametrize("kind", ["quad4", "quad4r", "tria3r"])
def test_fused_large_mesh_matches_oracle_and_twopass(kind):
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    case = cases.shell_mesh(kind, 61, 47, seed=404)
    want = driver.run(case, what=("KC0", "KG", "M0"))
    _check(case, want)
    b = util.batch_from_case(case)
    nn = case["ndof"] // 6
    plan = AssemblyPlan("KC0", nn, [b])
    coo, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True)
    two = b.evaluate(KC0=True, KG=True, M=True)
    for m in ("KC0", "KG", "M"):
        assert util.block_relerr(coo[m].v.cpu().numpy(), two[m].v.cpu().numpy(), case["conn"].shape[0]) <= 1e-13
        ref_csr = AssemblyPlan(m, nn, [b]).assemble(two[m].v)
        assert np.abs((csr[m] - ref_csr).cpu().numpy()).max() <= 1e-12 * float(ref_csr.abs().max())
    # CSR only (no COO arrays written)
    _, csr2 = plan.evaluate_assemble(KC0=True, KG=True, M=True, write_coo=False)
    for m in ("KC0", "KG", "M"):
        assert torch.equal(csr2[m], csr[m])


def _fan_case(kind, nquads=6, seed=7):
    """nquads quads around one centre node: the centre has valence nquads (> 4 -> two rounds of node records)."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(0, 2 * np.pi, 2 * nquads, endpoint=False)
    rim = np.stack([np.cos(ang), np.sin(ang), 0.05 * np.sin(3 * ang)], 1) * (1 + 0.1 * rng.uniform(-1, 1, (2 * nquads, 1)))
    X = np.vstack([[0., 0., 0.02], rim]) @ cases.random_rotation(rng).T
    conn = np.array([[0, 1 + 2 * i, 1 + (2 * i + 1) % (2 * nquads), 1 + (2 * i + 2) % (2 * nquads)]
                     for i in range(nquads)], np.int64)
    nn = X.shape[0]
    c = dict(kind=kind, x=X.ravel(), conn=conn, props=cases.random_shellprops(rng, 1), ndof=6 * nn,
             u=1e-4 * rng.normal(size=6 * nn), stress=(1e3, -2e2, 3e2))
    if kind == "quad4r":
        c["hg"] = rng.uniform(0.01, 1., (nquads, 5))
    return c


@pytest.mark.parametrize("kind", ["quad4", "quad4r"])
def test_fused_high_valence_node(kind):
    case = _fan_case(kind)
    want = driver.run(case, what=("KC0", "KG", "M0"))
    _check(case, want)


def _tria_fan_case(ntri=13, seed=9):
    """ntri triangles around one centre node: valence 13 > 10 -> two rounds of triangle node records."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(0, 2 * np.pi, ntri, endpoint=False)
    rim = np.stack([np.cos(ang), np.sin(ang), 0.05 * np.sin(3 * ang)], 1) * (1 + 0.1 * rng.uniform(-1, 1, (ntri, 1)))
    X = np.vstack([[0., 0., 0.02], rim]) @ cases.random_rotation(rng).T
    conn = np.array([[0, 1 + i, 1 + (i + 1) % ntri] for i in range(ntri)], np.int64)
    nn = X.shape[0]
    return dict(kind="tria3r", x=X.ravel(), conn=conn, props=cases.random_shellprops(rng, 1), ndof=6 * nn,
                u=1e-4 * rng.normal(size=6 * nn), stress=(1e3, -2e2, 3e2))


def test_fused_tria_high_valence_node():
    case = _tria_fan_case()
    want = driver.run(case, what=("KC0", "KG", "KGs", "M0"))
    _check(case, want)
    _check(case, want, stress=case["stress"])


def test_evaluate_assemble_falls_back_for_other_kinds():
    """BeamC has no fused kernel: the same call runs evaluation + slab assembly and returns the same outputs."""
    import scipy.sparse as sp
    from pyfe3d_b200.batch import AssemblyPlan
    case, ref = util.load_golden("beamc_chain")
    b = util.batch_from_case(case)
    n = case["ndof"]
    plan = AssemblyPlan("KC0", n // 6, [b])
    coo, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True, mtype=1, indices=True)
    for name, rk in (("KC0", "KC0"), ("KG", "KG"), ("M", "M1")):
        r, c, v = ref[rk]
        assert np.array_equal(coo[name].r.cpu().numpy(), r)
        assert util.block_relerr(coo[name].v.cpu().numpy(), v, case["conn"].shape[0]) <= util.TOL_VALUES
        A = plan._sibling(name, 1).to_scipy(csr[name])
        S = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
        S.sum_duplicates()
        assert abs(A - S).max() <= util.TOL_CSR * np.abs(S.data).max()


@pytest.mark.parametrize("kind", ["quad4", "tria3r"])
def test_host_buffer_step_matches_device_step(kind):
    """pf3_eval_assemble_host (x, u in and CSR values out through HOST buffers) == the device-resident call."""
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    case = cases.shell_mesh(kind, 17, 13, seed=21)
    b = util.batch_from_case(case)
    nn = case["ndof"] // 6
    plan = AssemblyPlan("KC0", nn, [b])
    _, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True, mtype=1, write_coo=False)
    sizes = plan.csr_sizes(1)
    out = {m: torch.empty(sizes[m], dtype=torch.float64).pin_memory() for m in ("KC0", "KG", "M")}
    xh = torch.as_tensor(np.ascontiguousarray(case["x"])).pin_memory()
    uh = torch.as_tensor(np.ascontiguousarray(case["u"])).pin_memory()
    plan.evaluate_assemble_host(xh, uh, out, KC0=True, KG=True, M=True, mtype=1)
    for m in ("KC0", "KG", "M"):
        assert torch.equal(out[m], csr[m].cpu()), m
    # a new displacement field through the host path changes KG only
    out2 = {m: np.empty(sizes[m]) for m in ("KC0", "KG")}
    plan.evaluate_assemble_host(case["x"], 3.0 * case["u"], out2, KC0=True, KG=True)
    assert np.array_equal(out2["KC0"], csr["KC0"].cpu().numpy())
    assert np.abs(out2["KG"] - 3.0 * csr["KG"].cpu().numpy()).max() <= 1e-12 * float(csr["KG"].abs().max()) * 3
    with pytest.raises(ValueError):
        plan.evaluate_assemble_host(torch.as_tensor(case["x"]).cuda(), uh, out, KC0=True)
