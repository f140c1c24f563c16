"""GPU: piston-theory aerodynamic matrices KA_beta / KA_gamma / CA of Quad4 and Quad4R --
batched kernel through the C ABI against the reference-generated fixtures (tests/golden/aero_*.npz), their CSR
assembly against scipy, and the per-element drop-in methods in the loop style of
tests/test_quad4r_piston_theory.py:84-124 of the reference."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
AERO = ("KA_beta", "KA_gamma", "CA")
NAMES = ["aero_quad4_mesh", "aero_quad4_soup", "aero_quad4r_mesh", "aero_quad4r_soup"]


@pytest.mark.parametrize("name", NAMES)
def test_batched_aero_matches_reference_golden(name):
    case, ref = util.load_golden(name)
    b = util.batch_from_case(case)
    ne = case["conn"].shape[0]
    coo = b.evaluate_aero(KA_beta=True, KA_gamma=True, CA=True)
    for k in AERO:
        r, c, v = ref[k]
        assert np.array_equal(coo[k].r.cpu().numpy(), r) and np.array_equal(coo[k].c.cpu().numpy(), c), k
        assert util.block_relerr(coo[k].v.cpu().numpy(), v, ne) <= util.TOL_VALUES, k
    # one matrix at a time, values only, gives the same values
    one = b.update_KA_beta(indices=False)
    assert one.r is None and np.array_equal(one.v.cpu().numpy(), coo["KA_beta"].v.cpu().numpy())


@pytest.mark.parametrize("name", ["aero_quad4_mesh", "aero_quad4r_mesh"])
def test_aero_assembly_matches_scipy(name):
    import scipy.sparse as sp
    from pyfe3d_b200.batch import AssemblyPlan
    case, ref = util.load_golden(name)
    b = util.batch_from_case(case)
    n = case["ndof"]
    coo = b.evaluate_aero(KA_beta=True, KA_gamma=True, CA=True, indices=False)
    for k in AERO:
        plan = AssemblyPlan(k, n // 6, [b])
        A = plan.to_scipy(plan.assemble(coo[k].v))
        r, c, v = ref[k]
        S = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
        S.sum_duplicates()
        S.sort_indices()
        assert np.array_equal(A.indptr, S.indptr) and np.array_equal(A.indices, S.indices), k
        assert np.abs(A.data - S.data).max() <= util.TOL_CSR * np.abs(S.data).max(), k


@pytest.mark.parametrize("name", ["aero_quad4_soup", "aero_quad4r_soup"])
def test_per_element_aero_methods(name):
    import pyfe3d_b200 as pf
    case, ref = util.load_golden(name)
    kind = case["kind"]
    cls = {"quad4": "Quad4", "quad4r": "Quad4R"}[kind]
    data = getattr(pf, cls + "Data")()
    probe = getattr(pf, cls + "Probe")()
    assert data.KA_BETA_SPARSE_SIZE == data.KA_GAMMA_SPARSE_SIZE == data.CA_SPARSE_SIZE == 144
    ne = 5
    conn = case["conn"][:ne]
    x = np.ascontiguousarray(case["x"], float)
    out = {k: [np.zeros(144 * ne, pf.INT), np.zeros(144 * ne, pf.INT), np.zeros(144 * ne)] for k in AERO}
    for e in range(ne):
        q = getattr(pf, cls)(probe)
        q.n1, q.n2, q.n3, q.n4 = [int(t) for t in conn[e]]
        q.c1, q.c2, q.c3, q.c4 = [6 * int(t) for t in conn[e]]
        q.init_k_KA_beta = q.init_k_KA_gamma = q.init_k_CA = 144 * e
        q.update_rotation_matrix(x)
        q.update_probe_xe(x)
        q.update_KA_beta(*out["KA_beta"])
        q.update_KA_gamma(*out["KA_gamma"])
        q.update_CA(*out["CA"])
    for k in AERO:
        r, c, v = [t[:144 * ne] for t in ref[k]]
        assert np.array_equal(out[k][0], r) and np.array_equal(out[k][1], c), k
        assert util.block_relerr(out[k][2], v, ne) <= util.TOL_VALUES, k
    # values accumulate like the reference's `+=` (quad4.pyx:9687)
    q.update_CA(*out["CA"])
    assert np.allclose(out["CA"][2][-144:], 2 * ref["CA"][2][144 * (ne - 1):144 * ne], rtol=1e-12, atol=0)
