"""GPU parity (run on the B200 box): the CUDA path, called through the C ABI, against
(a) the reference's golden vectors and (b) the numpy oracle on fresh seeded inputs.
Tolerances are north_star's: indices bit-exact, values 1e-12 (block-relative),
assembled CSR 1e-11."""
import numpy as np
import pytest

from oracle import coo as ocoo
from oracle import driver
from tests import cases, util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", util.golden_names())
def test_cuda_matches_reference_golden(name, fused):
    case, ref = util.load_golden(name)
    got = util.run_gpu(case, fused=fused)
    checked = util.compare_outputs(got, ref, case["conn"].shape[0])
    assert "KC0" in checked and "fint" in checked


@pytest.mark.parametrize("kind", cases.SHELL_KINDS + cases.LINE_KINDS)
def test_cuda_matches_oracle_fresh_inputs(kind):
    if kind in cases.SHELL_KINDS:
        case = cases.shell_soup(kind, 1000, seed=301)
    else:
        case = cases.line_soup(kind, 1000, seed=302, logL=(-0.5, 0.5))
    want = driver.run(case)
    got = util.run_gpu(case)
    util.compare_outputs(got, want, case["conn"].shape[0])


@pytest.mark.parametrize("kind", cases.SHELL_KINDS)
def test_cuda_mesh_matches_oracle(kind):
    case = cases.shell_mesh(kind, 23, 17, seed=303)
    want = driver.run(case)
    got = util.run_gpu(case)
    util.compare_outputs(got, want, case["conn"].shape[0])


def test_values_only_and_accumulate():
    import torch
    case, ref = util.load_golden("quad4_mesh")
    b = util.batch_from_case(case)
    k1 = b.update_KC0()
    k2 = b.update_KC0(update_KC0v_only=1)
    assert k2.r is None and k2.c is None
    assert torch.equal(k1.v, k2.v)
    # accumulate=True reproduces the reference's `+=` (quad4.pyx:1313)
    b.evaluate(KC0=True, out={"KC0": k1}, accumulate=True)
    # (the `+=` path runs the thread-per-element kernel, the overwrite path the pair-lane kernel: same
    # mathematics, different rounding order)
    assert util.block_relerr(k1.v.cpu().numpy(), 2 * k2.v.cpu().numpy(), case["conn"].shape[0]) <= 1e-13


@pytest.mark.parametrize("name", ["quad4_mesh", "quad4r_mesh", "tria3r_mesh", "beamc_chain", "truss_chain",
                                  "spring_chain", "beamlr_chain"])
@pytest.mark.parametrize("matrix", ["KC0", "KG", "M"])
def test_structured_assembly_matches_scipy(name, matrix):
    import scipy.sparse as sp
    from pyfe3d_b200.batch import AssemblyPlan
    case, ref = util.load_golden(name)
    b = util.batch_from_case(case)
    if b.sizes[matrix] == 0:
        pytest.skip("no such matrix")
    n = case["ndof"]
    key = {"KC0": "KC0", "KG": "KG", "M": "M0"}[matrix]
    coo = b.evaluate(**{matrix: True})[matrix]
    plan = AssemblyPlan(matrix, n // 6, [b])
    vals = plan.assemble(coo.v)
    A = plan.to_scipy(vals)
    r, c, v = ref[key]
    S = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
    S.sum_duplicates()
    S.sort_indices()
    np.testing.assert_array_equal(A.indptr, S.indptr)
    np.testing.assert_array_equal(A.indices, S.indices)
    scale = np.abs(S.data).max()
    assert np.abs(A.data - S.data).max() <= util.TOL_CSR * scale
    # generic (sort-based) plan on the same triplets
    from pyfe3d_b200.batch import CooPlan
    gp = CooPlan(n, coo.r, coo.c)
    G = gp.to_scipy(gp.assemble(coo.v))
    np.testing.assert_array_equal(G.indptr, S.indptr)
    np.testing.assert_array_equal(G.indices, S.indices)
    assert np.abs(G.data - S.data).max() <= util.TOL_CSR * scale


def test_lumped_mass_assembly_and_row_shards():
    """mtype 2 (tail entries stay (0,0,0)) and node-range shards reproduce the full matrix rows."""
    import scipy.sparse as sp
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    case, ref = util.load_golden("quad4_mesh")
    b = util.batch_from_case(case)
    n = case["ndof"]
    nn = n // 6
    coo = b.update_M(mtype=2)
    r, c, v = ref["M2"]
    S = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
    S.sum_duplicates()
    full = AssemblyPlan("M", nn, [b], mtype=2)
    A = full.to_scipy(full.assemble(coo.v))
    assert abs(A - S).max() <= util.TOL_CSR * np.abs(S.data).max()
    cut = nn // 3
    parts = []
    for lo, hi in ((0, cut), (cut, nn)):
        p = AssemblyPlan("M", nn, [b], node_range=(lo, hi), mtype=2)
        parts.append(p.to_scipy(p.assemble(coo.v)))
    stacked = sp.vstack(parts).tocsr()
    assert abs(stacked - A).max() == 0.0


def test_mixed_quad_beam_assembly():
    """Stiffened-panel style: Quad4 skin + BeamC stiffeners sharing nodes in ONE KC0 matrix."""
    import scipy.sparse as sp
    import torch
    from pyfe3d_b200.batch import AssemblyPlan, ElementBatch
    case = cases.shell_mesh("quad4", 9, 7, seed=77, curved=False)
    nn = case["ndof"] // 6
    rng = np.random.default_rng(5)
    pos = np.arange(9 * 7).reshape(9, 7)
    bconn = np.stack([pos[:-1, 3], pos[1:, 3]], 1).astype(np.int64)
    bcase = dict(kind="beamc", x=case["x"], conn=bconn, props=cases.random_beamprops(rng, 1),
                 vxy=np.tile([0.3, 0.2, 1.0], (bconn.shape[0], 1)), ndof=case["ndof"], u=case["u"])
    q = util.batch_from_case(case)
    bm = util.batch_from_case(bcase)
    kq, kb = q.update_KC0(), bm.update_KC0()
    v = torch.cat([kq.v, kb.v])
    plan = AssemblyPlan("KC0", nn, [q, bm])
    A = plan.to_scipy(plan.assemble(v))
    wq, wb = driver.run(case, what=("KC0",)), driver.run(bcase, what=("KC0",))
    r = np.concatenate([wq["KC0"][0], wb["KC0"][0]])
    c = np.concatenate([wq["KC0"][1], wb["KC0"][1]])
    vv = np.concatenate([wq["KC0"][2], wb["KC0"][2]])
    n = case["ndof"]
    S = sp.coo_matrix((vv, (r, c)), shape=(n, n)).tocsr()
    S.sum_duplicates()
    assert abs(A - S).max() <= util.TOL_CSR * np.abs(S.data).max()
    assert A.nnz == S.nnz


def test_spmv_matches_scipy():
    import torch
    from pyfe3d_b200.batch import AssemblyPlan, spmv
    case, _ = util.load_golden("quad4_mesh")
    b = util.batch_from_case(case)
    coo = b.update_KC0()
    plan = AssemblyPlan("KC0", case["ndof"] // 6, [b])
    vals = plan.assemble(coo.v)
    indptr, indices = plan.pattern()
    x = torch.as_tensor(case["u"]).cuda()
    y = spmv(indptr, indices, vals, x).cpu().numpy()
    want = plan.to_scipy(vals) @ case["u"]
    assert np.abs(y - want).max() <= 1e-12 * np.abs(want).max()


def test_state_and_finte():
    case, ref = util.load_golden("quad4r_soup")
    b = util.batch_from_case(case)
    st = b.state().cpu().numpy()
    ne = case["conn"].shape[0]
    np.testing.assert_allclose(st[:, :9].reshape(ne, 3, 3), ref["R"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(st[:, 9:13].reshape(ne, 2, 2), ref["m"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(st[:, 13], ref["geo"], rtol=1e-13)
    np.testing.assert_allclose(st[:, 14:26].reshape(ne, 4, 3), ref["xe"], rtol=0, atol=1e-14 * np.abs(ref["xe"]).max())
    fe = b.finte().cpu().numpy()
    assert fe.shape == (ne, 24) and np.isfinite(fe).all()


def test_two_group_plan_with_odd_offset():
    """Two Tria3R batches in ONE KG matrix: group 1 starts at an odd COO offset (81 * odd), so its slabs are
    only 8-byte aligned (the slab assembly must not use 16-byte async copies there)."""
    import scipy.sparse as sp
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    case = cases.shell_mesh("tria3r", 6, 5, seed=91)
    ne = case["conn"].shape[0]
    cut = 7
    parts = []
    for sel in (slice(0, cut), slice(cut, ne)):
        c = dict(case)
        c["conn"] = case["conn"][sel]
        c["prop_id"] = case["prop_id"][sel]
        c["xmat"] = case["xmat"][sel]
        parts.append(c)
    bs = [util.batch_from_case(c) for c in parts]
    coos = [b.update_KG() for b in bs]
    v = torch.cat([c.v for c in coos])
    nn = case["ndof"] // 6
    plan = AssemblyPlan("KG", nn, bs)
    assert plan.coo_offsets[1] % 2 == 1
    A = plan.to_scipy(plan.assemble(v))
    want = driver.run(case, what=("KG",))["KG"]
    n = case["ndof"]
    S = sp.coo_matrix((want[2], (want[0], want[1])), shape=(n, n)).tocsr()
    S.sum_duplicates()
    assert abs(A - S).max() <= util.TOL_CSR * np.abs(S.data).max()
    assert A.nnz == S.nnz


@pytest.mark.parametrize("name", ["quad4_mesh", "tria3r_mesh", "beamc_chain", "spring_chain"])
def test_plan_fint_matches_reference(name):
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    case, ref = util.load_golden(name)
    b = util.batch_from_case(case)
    nn = case["ndof"] // 6
    plan = AssemblyPlan("KC0", nn, [b])
    fint = torch.zeros(case["ndof"], dtype=torch.float64, device=b.device)
    plan.update_fint(fint)
    assert util.vec_relerr(fint.cpu().numpy(), ref["fint"]) <= util.TOL_VALUES
    # row shard: only owned rows are touched
    lo, hi = nn // 4, nn - 1
    ps = AssemblyPlan("KC0", nn, [b], node_range=(lo, hi))
    f2 = torch.zeros(case["ndof"], dtype=torch.float64, device=b.device)
    ps.update_fint(f2)
    f2 = f2.cpu().numpy()
    assert np.all(f2[:6 * lo] == 0) and np.all(f2[6 * hi:] == 0)
    assert util.vec_relerr(f2[6 * lo:6 * hi], ref["fint"][6 * lo:6 * hi]) <= util.TOL_VALUES * (
        np.abs(ref["fint"]).max() / max(np.abs(ref["fint"][6 * lo:6 * hi]).max(), 1e-300))
