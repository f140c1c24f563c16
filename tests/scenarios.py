"""Analysis scenarios in the style of the reference's own test scripts, beyond the five BASELINE configs of
tests/configs.py: two-stage linear buckling (static solve -> KG from the displacements -> eigenproblem), pre-stressed
natural frequencies, buckling under a given stress state, frequencies with consistent / lumped mass for every line
element, a mixed truss + spring model.  Each scenario is ONE function of an ``evaluate(case, keys)`` callback returning
{key: scipy CSR}; the same function runs with matrices from the compiled reference (tests/golden/make_scenario_scalars.py
-> tests/golden/scenario_scalars.json), from the numpy oracle (CPU test) and from the CUDA path (GPU test), so the
comparison at north_star's 1e-8 isolates element evaluation + assembly.  All eigenproblems are solved densely
(scipy.linalg.eigh) so that no iterative-solver tolerance enters the comparison.

Reference scripts the scenarios follow (geometry, materials, boundary conditions):
  quad4_plate_buckling, quad4r_plate_buckling   tests/test_quad4r_linear_buckling_plate.py:20-190
  tria3r_buckling_given_stress                  tests/test_tria3r_linear_buckling_plate_given_stress.py
  quad4r_prestress_freq                         tests/test_quad4r_natural_freq_pre_stress.py
  beamc_prestress_freq                          tests/test_beamc_natural_freq_cantilever_pre_stress.py
  beamlr_cantilever_freq                        tests/test_beamlr_natural_freq_cantilever.py:15-111
  truss_freq                                    tests/test_truss_natural_freq.py:13-104
  quad4_freq_mat_coord                          tests/test_quad4_natural_freq_distorted_mat_coord.py
"""
import numpy as np
import scipy.linalg as sl
from scipy.sparse.linalg import spsolve

from pyfe3d_b200 import meshes
from pyfe3d_b200.shellprop_utils import isotropic_plate, laminated_plate

DOF = 6
NAMES = ("quad4_plate_buckling", "quad4r_plate_buckling", "tria3r_buckling_given_stress", "quad4r_prestress_freq",
         "beamc_prestress_freq", "beamlr_cantilever_freq", "truss_freq", "quad4_freq_mat_coord")


# ---------------------------------------------------------------------------------------------------- helpers
def _plate(kind, nx, ny, a, b, prop, rotate=True, distort=0., seed=3):
    """(nx x ny)-element plate on [0,a]x[0,b]; returns the case and the plate-frame node coordinates."""
    rng = np.random.default_rng(seed)
    nnx, nny = nx + 1, ny + 1
    xs, ys = np.meshgrid(np.linspace(0, a, nnx), np.linspace(0, b, nny), indexing="ij")
    if distort:
        xs[1:-1, 1:-1] += distort * a / nx * (2 * rng.random((nnx - 2, nny - 2)) - 1)
        ys[1:-1, 1:-1] += distort * b / ny * (2 * rng.random((nnx - 2, nny - 2)) - 1)
    Xl = np.stack([xs.ravel(), ys.ravel(), np.zeros(nnx * nny)], 1)
    Q = meshes.fixed_rotation(seed) if rotate else np.eye(3)
    X = Xl @ Q.T
    pos = np.arange(nnx * nny).reshape(nnx, nny)
    n1, n2, n3, n4 = pos[:-1, :-1].ravel(), pos[1:, :-1].ravel(), pos[1:, 1:].ravel(), pos[:-1, 1:].ravel()
    if kind == "tria3r":
        conn = np.concatenate([np.stack([n1, n2, n3], 1), np.stack([n1, n3, n4], 1)])
    else:
        conn = np.stack([n1, n2, n3, n4], 1)
    case = dict(kind=kind, x=X.ravel(), conn=conn.astype(np.int64), props=meshes.shellprop_row(prop)[None, :],
                ndof=DOF * nnx * nny, u=np.zeros(DOF * nnx * nny))
    return case, Xl, Q


def _line(kind, n, L, props, vxy=(0., 0., 1.), seed=4, rotate=True):
    Q = meshes.fixed_rotation(seed) if rotate else np.eye(3)
    Xl = np.stack([np.linspace(0, L, n), np.zeros(n), np.zeros(n)], 1)
    conn = np.stack([np.arange(n - 1), np.arange(1, n)], 1).astype(np.int64)
    case = dict(kind=kind, x=(Xl @ Q.T).ravel(), conn=conn, props=props, ndof=DOF * n, u=np.zeros(DOF * n))
    if kind != "truss":
        case["vxy"] = np.tile(Q @ np.asarray(vxy), (n - 1, 1))
    return case, Xl, Q


def _beamprops(E, nu, rho, by, hz):
    A, Izz, Iyy = by * hz, hz * by ** 3 / 12, by * hz ** 3 / 12
    p = np.zeros((1, 16))
    p[0, :9] = [A, E, E / 2 / (1 + nu) * 5 / 6., Iyy, Izz, 0., Iyy + Izz, 0., 0.]
    p[0, 9:15] = [rho * A, 0., 0., rho * Izz, rho * Iyy, 0.]
    return p


def _sub(A, bu):
    return A.tocsc()[bu, :][:, bu]


def _solve(K, f, bu):
    u = np.zeros(f.size)
    u[bu] = spsolve(_sub(K, bu), f[bu])
    return u


def _buckling(K, KG, bu, k=2):
    """smallest positive load factors of (K + lambda KG) phi = 0, dense."""
    Kd, Gd = _sub(K, bu).toarray(), _sub(KG, bu).toarray()
    mu = sl.eigh(-0.5 * (Gd + Gd.T), 0.5 * (Kd + Kd.T), eigvals_only=True)   # mu = 1 / lambda
    lam = np.sort(1. / mu[mu > 1e-14 * np.abs(mu).max()])
    return lam[:k]


def _freqs(K, M, bu, k=3):
    """lowest natural frequencies, dense: mu = 1 / omega^2 from eigh(M, K) -- K is positive definite on the free DOFs
    while M may be singular (no drilling inertia, lumped matrices), so K is the matrix that gets factorised."""
    Kd, Md = _sub(K, bu).toarray(), _sub(M, bu).toarray()
    mu = sl.eigh(0.5 * (Md + Md.T), 0.5 * (Kd + Kd.T), eigvals_only=True)
    return 1. / np.sqrt(np.sort(mu)[::-1][:k])


def _ss_plate_bcs(Xl, a, b, Q):
    """Simply supported edges: the plate-normal translation is held on the boundary.  With a rotated plate the
    constraint is not axis aligned, so the scenarios hold ALL translations on the edges and release rotations."""
    x, y = Xl[:, 0], Xl[:, 1]
    edge = np.isclose(x, 0.) | np.isclose(x, a) | np.isclose(y, 0.) | np.isclose(y, b)
    bk = np.zeros(DOF * Xl.shape[0], bool)
    for d in range(3):
        bk[d::DOF] = edge
    return bk, edge


# ---------------------------------------------------------------------------------------------------- scenarios
def _shell_two_stage_buckling(evaluate, kind):
    a, b, nx, ny = 0.8, 0.5, 10, 8
    prop = laminated_plate(stack=[0, 45, -45, 90, 90, -45, 45, 0], plyt=0.125e-3,
                           laminaprop=(142.5e9, 8.7e9, 0.28, 5.1e9, 5.1e9, 5.1e9))
    case, Xl, Q = _plate(kind, nx, ny, a, b, prop, rotate=True, distort=0.15)
    nn = Xl.shape[0]
    x, y = Xl[:, 0], Xl[:, 1]
    # stage 1: uniaxial compression along the plate x axis; x = 0 held in plate-x, the normal held on all edges,
    # one corner held in plate-y.  Constraints are expressed in the plate frame through multipoint-free penalty-less
    # elimination: rotate the problem back (K_local = T^T K T) so that axis-aligned BCs can be applied.
    T = np.kron(np.eye(2 * nn), Q)                     # global = T @ local, per 3-vector
    K = evaluate(case, ("KC0",))["KC0"]
    Kl = (T.T @ K.toarray() @ T)
    edge = np.isclose(x, 0.) | np.isclose(x, a) | np.isclose(y, 0.) | np.isclose(y, b)
    bk = np.zeros(DOF * nn, bool)
    bk[2::DOF] = edge
    bk[0::DOF] = np.isclose(x, 0.)
    bk[1::DOF] = np.isclose(x, 0.) & np.isclose(y, 0.)
    bk[5::DOF] = np.isclose(x, 0.) & np.isclose(y, 0.)
    bu = ~bk
    f = np.zeros(DOF * nn)
    right = np.isclose(x, a)
    w = np.where(np.isclose(y, 0.) | np.isclose(y, b), 0.5, 1.0) * (b / ny)
    f[0::DOF][right] = -1000. * w[right]               # N/m line load, lumped to the nodes
    ul = np.zeros(DOF * nn)
    ul[bu] = np.linalg.solve(Kl[np.ix_(bu, bu)], f[bu])
    u = T @ ul
    # stage 2: KG from the displacement field, buckling factors in the plate frame
    KG = evaluate(dict(case, u=u), ("KG",))["KG"]
    Gl = T.T @ KG.toarray() @ T
    Kd, Gd = Kl[np.ix_(bu, bu)], Gl[np.ix_(bu, bu)]
    mu = sl.eigh(-0.5 * (Gd + Gd.T), 0.5 * (Kd + Kd.T), eigvals_only=True)
    lam = np.sort(1. / mu[mu > 1e-14 * np.abs(mu).max()])
    return {"u_norm": float(np.linalg.norm(u)), "lambda1": float(lam[0]), "lambda2": float(lam[1])}


def quad4_plate_buckling(evaluate):
    return _shell_two_stage_buckling(evaluate, "quad4")


def quad4r_plate_buckling(evaluate):
    return _shell_two_stage_buckling(evaluate, "quad4r")


def tria3r_buckling_given_stress(evaluate):
    a, b = 0.6, 0.4
    prop = isotropic_plate(thickness=0.002, E=70e9, nu=0.33, rho=2700.)
    case, Xl, Q = _plate("tria3r", 9, 7, a, b, prop, rotate=True, distort=0.2)
    case["stress"] = (-100., 0., -30.)
    bk, _ = _ss_plate_bcs(Xl, a, b, Q)
    mats = evaluate(case, ("KC0", "KGs"))
    lam = _buckling(mats["KC0"], mats["KGs"], ~bk)
    return {"lambda1": float(lam[0]), "lambda2": float(lam[1])}


def quad4r_prestress_freq(evaluate):
    a, b, nx, ny = 0.5, 0.4, 9, 8
    prop = isotropic_plate(thickness=0.003, E=70e9, nu=0.33, rho=2700.)
    case, Xl, Q = _plate("quad4r", nx, ny, a, b, prop, rotate=False, distort=0.2)
    nn = Xl.shape[0]
    x, y = Xl[:, 0], Xl[:, 1]
    edge = np.isclose(x, 0.) | np.isclose(x, a) | np.isclose(y, 0.) | np.isclose(y, b)
    bk = np.zeros(DOF * nn, bool)
    bk[2::DOF] = edge
    bk[0::DOF] = np.isclose(x, 0.)
    bk[1::DOF] = np.isclose(x, 0.) & np.isclose(y, 0.)
    bu = ~bk
    # in-plane tension along x stiffens the bending modes
    f = np.zeros(DOF * nn)
    right = np.isclose(x, a)
    w = np.where(np.isclose(y, 0.) | np.isclose(y, b), 0.5, 1.0) * (b / ny)
    f[0::DOF][right] = 2.0e5 * w[right]
    K = evaluate(case, ("KC0",))["KC0"]
    u = _solve(K, f, bu)
    mats = evaluate(dict(case, u=u), ("KG", "M0"))
    om0 = _freqs(K, mats["M0"], bu, 2)
    om1 = _freqs(K + mats["KG"], mats["M0"], bu, 2)
    return {"u_norm": float(np.linalg.norm(u)), "omega1": float(om0[0]), "omega1_prestress": float(om1[0]),
            "omega2_prestress": float(om1[1])}


def beamc_prestress_freq(evaluate):
    E, nu, rho, L, n = 203e9, 0.3, 7830., 3., 41
    case, Xl, Q = _line("beamc", n, L, _beamprops(E, nu, rho, 0.05, 0.07), vxy=(0., 1., 0.))
    bk = np.zeros(DOF * n, bool)
    bk[:DOF] = True
    bu = ~bk
    f = np.zeros(DOF * n)
    f[DOF * (n - 1):DOF * (n - 1) + 3] = Q @ np.array([-1.5e4, 0., 0.])      # axial compression at the tip (0.37 of the Euler load)
    K = evaluate(case, ("KC0",))["KC0"]
    u = _solve(K, f, bu)
    mats = evaluate(dict(case, u=u), ("KG", "M0", "M1"))
    om = _freqs(K, mats["M0"], bu, 2)
    omp = _freqs(K + mats["KG"], mats["M0"], bu, 2)
    oml = _freqs(K + mats["KG"], mats["M1"], bu, 1)
    return {"tip": float(np.linalg.norm(u[-DOF:-3])), "omega1": float(om[0]), "omega2": float(om[1]),
            "omega1_prestress": float(omp[0]), "omega2_prestress": float(omp[1]), "omega1_prestress_lumped": float(oml[0])}


def beamlr_cantilever_freq(evaluate):
    E, nu, rho, L, n = 203e9, 0.3, 7830., 3., 61
    case, Xl, Q = _line("beamlr", n, L, _beamprops(E, nu, rho, 0.05, 0.04), vxy=(0., 1., 0.))
    bk = np.zeros(DOF * n, bool)
    bk[:DOF] = True
    mats = evaluate(case, ("KC0", "M0", "M1"))
    om = _freqs(mats["KC0"], mats["M0"], ~bk, 3)
    oml = _freqs(mats["KC0"], mats["M1"], ~bk, 2)
    A, Imin = 0.05 * 0.04, 0.05 * 0.04 ** 3 / 12
    euler = 1.875 ** 2 * np.sqrt(E * Imin / (rho * A * L ** 4))              # the reference's analytic check
    return {"omega1": float(om[0]), "omega2": float(om[1]), "omega3": float(om[2]), "omega1_lumped": float(oml[0]),
            "omega2_lumped": float(oml[1]), "omega1_over_euler": float(om[0] / euler)}


def truss_freq(evaluate):
    E, nu, rho, L, n = 203e9, 0.3, 7830., 3., 51
    case, Xl, Q = _line("truss", n, L, _beamprops(E, nu, rho, 0.05, 0.05), rotate=False)
    bk = np.ones(DOF * n, bool)
    bk[0::DOF] = False                                  # only the axial translation is free ...
    bk[0] = True                                        # ... and fixed at x = 0
    mats = evaluate(case, ("KC0", "M0", "M1"))
    om = _freqs(mats["KC0"], mats["M0"], ~bk, 3)
    oml = _freqs(mats["KC0"], mats["M1"], ~bk, 2)
    exact = np.pi / L / 2 * np.sqrt(E / rho)
    return {"omega1": float(om[0]), "omega2": float(om[1]), "omega3": float(om[2]), "omega1_lumped": float(oml[0]),
            "omega2_lumped": float(oml[1]), "omega1_over_exact": float(om[0] / exact)}


def quad4_freq_mat_coord(evaluate):
    a, b = 0.6, 0.45
    prop = laminated_plate(stack=[30, -30, 0, 0, -30, 30], plyt=0.2e-3,
                           laminaprop=(142.5e9, 8.7e9, 0.28, 5.1e9, 5.1e9, 5.1e9), rho=1600.)
    case, Xl, Q = _plate("quad4", 9, 8, a, b, prop, rotate=True, distort=0.25)
    ne = case["conn"].shape[0]
    case["xmat"] = np.tile(Q @ np.array([np.cos(0.4), np.sin(0.4), 0.]), (ne, 1))   # material axis 0.4 rad off plate-x
    bk, _ = _ss_plate_bcs(Xl, a, b, Q)
    mats = evaluate(case, ("KC0", "M0", "M2"))
    om = _freqs(mats["KC0"], mats["M0"], ~bk, 3)
    oml = _freqs(mats["KC0"], mats["M2"], ~bk, 2)
    return {"omega1": float(om[0]), "omega2": float(om[1]), "omega3": float(om[2]), "omega1_lumped": float(oml[0]),
            "omega2_lumped": float(oml[1])}


SCENARIOS = {n: globals()[n] for n in NAMES}


# ---------------------------------------------------------------------------------------------------- evaluators
def evaluate_with(run):
    """evaluate(case, keys) from a ``run(case, what)`` that returns COO triplets (oracle.driver.run, ref_loop.run)."""
    import scipy.sparse as sp

    def evaluate(case, keys):
        out = run(case, what=tuple(keys))
        n = case["ndof"]
        return {k: sp.coo_matrix((out[k][2], (out[k][0], out[k][1])), shape=(n, n)).tocsr() for k in keys}
    return evaluate


def evaluate_cuda(case, keys):
    """CUDA path: one fused evaluate+assemble call per scenario stage where the kind has one (Quad4/Quad4R/Tria3R),
    evaluation + structured assembly otherwise; CSR values come back through the plan's own pattern."""
    from pyfe3d_b200.batch import AssemblyPlan
    from tests import util
    b = util.batch_from_case(case)
    nn = case["ndof"] // DOF
    plan = AssemblyPlan("KC0", nn, [b])
    out = {}
    mkeys = [k for k in keys if k.startswith("M")]
    first = True
    for mk in (mkeys or [None]):
        mtype = int(mk[1]) if mk else 0
        kw = dict(KC0=first and "KC0" in keys, KG=first and "KG" in keys,
                  KG_given_stress=case.get("stress") if (first and "KGs" in keys) else None, M=mk is not None, mtype=mtype)
        _, csr = plan.evaluate_assemble(write_coo=False, **kw)
        for name, vals in csr.items():
            key = mk if name == "M" else ("KGs" if (name == "KG" and "KGs" in keys) else name)
            out[key] = plan._sibling(name, mtype if name == "M" else 0).to_scipy(vals)
        first = False
    return out
