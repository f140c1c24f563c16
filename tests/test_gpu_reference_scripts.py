"""GPU: user scripts written the way the reference's tests are written -- `import pyfe3d_b200 as pyfe3d`, one element
object per element, the reference's update_* calls, scipy for assembly and solve -- with the ANALYTIC assertions the
reference's own scripts make:
  truss tetrahedron under a vertical tip load     tests/test_truss_static.py:113-128   (w = P L / (2 A E), fint == fext)
  three springs, all six DOFs loaded              tests/test_spring.py:86-100          (series/parallel formulas, fint == fext)
  cantilever BeamC / BeamLR with a point load     tests/test_beamc_static_point_load.py:111-124 (7.48768 at rtol 1e-3)
"""
import numpy as np
import pytest
from scipy.sparse import coo_matrix
from scipy.sparse.linalg import spsolve

import pyfe3d_b200 as pyfe3d
from pyfe3d_b200 import DOF, DOUBLE, INT
from pyfe3d_b200.beamprop import BeamProp

pytestmark = pytest.mark.gpu


def _reactions_close(KC0, bk, bu, u, fext, fint):
    f = fext.copy()
    f[bk] = KC0[bk, :][:, bu] @ u[bu]      # reaction forces join the external vector (tests/test_truss_static.py:126)
    assert np.abs(fint - f).max() <= 1e-9 * np.abs(f).max()      # the reference's allclose, scaled to the load


def test_truss_tetrahedron_static():
    L, E, A = 3., 203.e9, 3.e-4
    x = np.array([0, L, L / 2, L / 2])
    y = np.array([0, 0, L * 3 ** 0.5 / 2, L * 3 ** 0.5 / 6])
    z = np.array([0, 0, 0, L * 6 ** 0.5 / 3])
    ncoords = np.ascontiguousarray(np.vstack((x, y, z)).T.flatten())
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    data, probe = pyfe3d.TrussData(), pyfe3d.TrussProbe()
    size = data.KC0_SPARSE_SIZE
    KC0r, KC0c = np.zeros(size * len(pairs), dtype=INT), np.zeros(size * len(pairs), dtype=INT)
    KC0v = np.zeros(size * len(pairs), dtype=DOUBLE)
    N = DOF * 4
    prop = BeamProp()
    prop.A, prop.E, prop.G = A, E, 5 / 6. * E / 2 / 1.3
    prop.intrho = 7.83e3 * A
    trusses = []
    for i, (p1, p2) in enumerate(pairs):
        t = pyfe3d.Truss(probe)
        t.init_k_KC0 = i * size
        t.n1, t.n2, t.c1, t.c2 = p1 + 1, p2 + 1, DOF * p1, DOF * p2
        t.update_rotation_matrix(ncoords)
        t.update_probe_xe(ncoords)
        t.update_KC0(KC0r, KC0c, KC0v, prop)
        trusses.append(t)
    KC0 = coo_matrix((KC0v, (KC0r, KC0c)), shape=(N, N)).tocsc()
    bk = np.zeros(N, dtype=bool)
    base, top = np.isclose(z, 0.), np.isclose(z, z.max())
    for d in range(3):
        bk[d::DOF][base] = True
    for d in range(3, 6):
        bk[d::DOF] = True
    bu = ~bk
    P = -7.
    fext = np.zeros(N)
    fext[2::DOF][top] = P
    u = np.zeros(N)
    u[bu] = spsolve(KC0[bu, :][:, bu], fext[bu])
    assert np.isclose(P * L / (2 * A * E), u[2::DOF].min())
    fint = np.zeros(N)
    for t in trusses:
        t.update_probe_ue(u)
        t.update_fint(fint, prop)
    _reactions_close(KC0, bk, bu, u, fext, fint)


def test_spring_chain_all_dofs():
    data, probe = pyfe3d.SpringData(), pyfe3d.SpringProbe()
    size = data.KC0_SPARSE_SIZE
    KC0r, KC0c, KC0v = np.zeros(size * 3, dtype=INT), np.zeros(size * 3, dtype=INT), np.zeros(size * 3, dtype=DOUBLE)
    N = DOF * 3
    k, kr = 1., 3.
    springs = []
    for i, (n1, n2) in enumerate([(0, 1), (0, 1), (1, 2)]):      # two springs in parallel, then one in series
        s = pyfe3d.Spring(probe)
        s.init_k_KC0 = i * size
        s.n1, s.n2, s.c1, s.c2 = n1, n2, n1 * DOF, n2 * DOF
        s.kxe = s.kye = s.kze = k
        s.krxe = s.krye = s.krze = kr
        s.update_rotation_matrix(1, 0, 0, 1, 1, 0)
        s.update_KC0(KC0r, KC0c, KC0v)
        springs.append(s)
    KC0 = coo_matrix((KC0v, (KC0r, KC0c)), shape=(N, N)).tocsc()
    bk = np.zeros(N, dtype=bool)
    bk[:DOF] = True
    bu = ~bk
    fext = np.zeros(N)
    fext[2 * DOF:] = [3., 5., 7., 11., 13., 17.]
    u = np.zeros(N)
    u[bu] = spsolve(KC0[bu, :][:, bu], fext[bu])
    stiff = np.array([k, k, k, kr, kr, kr])
    u2 = fext[2 * DOF:] / (2 * stiff)
    assert np.allclose(u2, u[DOF:2 * DOF])
    assert np.allclose(3 * u2, u[2 * DOF:])
    fint = np.zeros(N)
    for s in springs:
        s.update_probe_ue(u)
        s.update_fint(fint)
    _reactions_close(KC0, bk, bu, u, fext, fint)


@pytest.mark.parametrize("name,rtol", [("BeamC", 1e-3), ("BeamLR", 2e-3)])
def test_cantilever_point_load(name, rtol):
    n, L, a, P, E, nu = 33, 8., 5., -10.e3, 203.e9, 0.3
    x = np.linspace(0, L, n)
    hy = hz = 0.05
    ncoords = np.ascontiguousarray(np.vstack((x, np.ones_like(x), np.zeros_like(x))).T.flatten())
    data, probe = getattr(pyfe3d, name + "Data")(), getattr(pyfe3d, name + "Probe")()
    size = data.KC0_SPARSE_SIZE
    ne = n - 1
    KC0r, KC0c, KC0v = np.zeros(size * ne, dtype=INT), np.zeros(size * ne, dtype=INT), np.zeros(size * ne, dtype=DOUBLE)
    N = DOF * n
    prop = BeamProp()
    prop.A, prop.E = hy * hz, E
    prop.G = 5 / 6. * E / 2 / (1 + nu)
    prop.Izz, prop.Iyy = hz * hy ** 3 / 12, hy * hz ** 3 / 12
    prop.J = prop.Izz + prop.Iyy
    beams = []
    for i in range(ne):
        b = getattr(pyfe3d, name)(probe)
        b.init_k_KC0 = i * size
        b.n1, b.n2, b.c1, b.c2 = i + 1, i + 2, DOF * i, DOF * (i + 1)
        b.update_rotation_matrix(1., 1., 0., ncoords)
        b.update_probe_xe(ncoords)
        b.update_KC0(KC0r, KC0c, KC0v, prop)
        beams.append(b)
    KC0 = coo_matrix((KC0v, (KC0r, KC0c)), shape=(N, N)).tocsc()
    bk = np.zeros(N, dtype=bool)
    bk[:DOF] = True
    bu = ~bk
    fext = np.zeros(N)
    load = np.isclose(x, a)
    assert load.sum() == 1
    fext[1::DOF][load] = P
    fext[2::DOF][load] = -P
    u = np.zeros(N)
    u[bu] = spsolve(KC0[bu, :][:, bu], fext[bu])
    ref_value = 7.48768        # P a^2 (3 L - a) / (6 E I), tests/test_beamc_static_point_load.py:111-112
    assert np.isclose(-P * a ** 2 * (3 * L - a) / (6 * E * prop.Izz), ref_value, rtol=1e-5)
    assert np.isclose(u[1::DOF].min(), -ref_value, rtol=rtol)
    assert np.isclose(u[2::DOF].max(), +ref_value, rtol=rtol)
    fint = np.zeros(N)
    for b in beams:
        b.update_probe_ue(u)
        b.update_fint(fint, prop)
    _reactions_close(KC0, bk, bu, u, fext, fint)
