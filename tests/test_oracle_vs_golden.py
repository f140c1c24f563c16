"""CPU: pin the numpy oracle against the reference's golden vectors
(tests/golden/*.npz, made by tests/golden/make_golden.py from the compiled
reference) and, when oracle/_ref is present, against the reference run live."""
import numpy as np
import pytest

from oracle import coo, driver, ref_loop
from tests import cases, util


@pytest.mark.parametrize("name", util.golden_names())
def test_oracle_matches_golden(name):
    case, ref = util.load_golden(name)
    got = driver.run(case)
    checked = util.compare_outputs(got, ref, case["conn"].shape[0])
    assert "KC0" in checked and "fint" in checked
    assert set(k for k in got if k not in ("R", "m", "xe", "geo")) == set(
        k for k in ref if k not in ("R", "m", "xe", "geo"))


@pytest.mark.skipif(not ref_loop.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind", cases.SHELL_KINDS + cases.LINE_KINDS)
def test_oracle_matches_live_reference(kind):
    if kind in cases.SHELL_KINDS:
        case = cases.shell_soup(kind, 40, seed=101)
    else:
        case = cases.line_soup(kind, 40, seed=102)
    ref = ref_loop.run(case)
    got = driver.run(case)
    util.compare_outputs(got, ref, case["conn"].shape[0])


def test_golden_inputs_reproducible():
    """The committed fixtures were generated from tests/cases.py seeds."""
    gc = cases.golden_cases()
    for name in util.golden_names():
        case, _ = util.load_golden(name)
        np.testing.assert_array_equal(case["conn"], gc[name]["conn"])
        np.testing.assert_allclose(case["x"], gc[name]["x"], rtol=0, atol=0)


def test_coo_to_csr_matches_scipy():
    import scipy.sparse as sp
    case, ref = util.load_golden("quad4_mesh")
    r, c, v = ref["KC0"]
    n = case["ndof"]
    indptr, indices, data = coo.coo_to_csr(n, r, c, v)
    A = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    np.testing.assert_array_equal(indptr, A.indptr)
    np.testing.assert_array_equal(indices, A.indices)
    assert util.vec_relerr(data, A.data) < 1e-13


def test_quirk_tria3r_kc0_drops_drilling_couplings():
    """Tria3R update_KC0 omits (u,v)-rz couplings that update_fint keeps,
    so KC0 @ u differs from fint at ~1e-6 relative; Quad4 is self-consistent."""
    for kind, lo, hi in (("tria3r", 1e-9, 1e-3), ("quad4", 0.0, 1e-12)):
        case, ref = util.load_golden(kind + "_mesh")
        r, c, v = ref["KC0"]
        n = case["ndof"]
        Ku = np.zeros(n)
        np.add.at(Ku, r, v * case["u"][c])
        err = np.abs(Ku - ref["fint"]).max() / np.abs(ref["fint"]).max()
        assert lo <= err <= hi, (kind, err)


AERO = ("KA_beta", "KA_gamma", "CA")


@pytest.mark.parametrize("name", ["aero_quad4_mesh", "aero_quad4_soup", "aero_quad4r_mesh", "aero_quad4r_soup"])
def test_oracle_aero_matches_golden(name):
    """Piston-theory matrices: oracle vs the reference-generated fixtures."""
    case, ref = util.load_golden(name)
    got = driver.run(case, what=AERO)
    checked = util.compare_outputs(got, ref, case["conn"].shape[0], keys=AERO)
    assert set(checked) == set(AERO)
    # CA = -KA_gamma in the reference (quad4.pyx:11301 ff. against :10498 ff.)
    assert np.array_equal(ref["CA"][2], -ref["KA_gamma"][2])


@pytest.mark.skipif(not ref_loop.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind", ["quad4", "quad4r"])
def test_oracle_aero_matches_live_reference(kind):
    case = cases.shell_soup(kind, 30, seed=103)
    ref = ref_loop.run(case, what=AERO)
    got = driver.run(case, what=AERO)
    util.compare_outputs(got, ref, case["conn"].shape[0], keys=AERO)
