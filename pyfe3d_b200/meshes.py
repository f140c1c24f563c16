"""Synthetic structured meshes of the BASELINE.json configurations (numpy, host side).

Every generator returns a *case* dict (see oracle/driver.py for the fields) so the same
mesh can be handed to the CUDA path (``ElementBatch``), the numpy oracle and the compiled
reference.  Node numbering is i-major: node (i, j) has position i*(ny+1)+j, so a
contiguous range of node positions is a strip of the mesh (row-ownership shards)."""
import numpy as np

from .shellprop_utils import laminated_plate

FIELDS = ["A11", "A12", "A16", "A22", "A26", "A66", "B11", "B12", "B16", "B22", "B26", "B66",
          "D11", "D12", "D16", "D22", "D26", "D66", "E44", "E45", "E55", "scf_k13", "scf_k23", "h",
          "intrho", "intrhoz", "intrhoz2"]


def shellprop_row(prop):
    row = np.zeros(32)
    for j, f in enumerate(FIELDS):
        row[j] = getattr(prop, f)
    return row


def fixed_rotation(seed=0):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    return q


def north_star_laminate():
    """single coupled laminate [30,-45,0], plyt = 1e-3, offset so that B != 0."""
    return laminated_plate(stack=[30, -45, 0], plyt=1e-3,
                           laminaprop=(142.5e9, 8.7e9, 0.28, 5.1e9, 5.1e9, 5.1e9), rho=1600.,
                           offset=0.5e-3)


def plate_quad4(nx, ny, a=1.0, b=1.0, i0=None, i1=None, kind="quad4", seed=0, with_u=True, local=False):
    """nx x ny Quad4 elements on an a x b plate, rigidly rotated (R != I), u = 1e-4 N(0,1).

    ``i0, i1``: return only the elements that touch node columns [i0, i1) ("halo elements
    duplicated").  With ``local=False`` node arrays stay global; with ``local=True`` only the
    node columns those elements touch are generated and numbered from 0 (what one rank of a row-ownership
    shard holds: memory and host<->device traffic independent of the number of ranks); global node
    position = local + ``node_offset``.  Returns the case plus ``owned_nodes = (begin, end)`` and
    ``owned_elements`` (elements whose first node is owned: a disjoint cover used to count unique elements)."""
    nnx, nny = nx + 1, ny + 1
    lo = 0 if i0 is None else max(i0 - 1, 0)
    hi = nx if i1 is None else min(i1, nx)
    c0, c1 = (lo, hi) if local else (0, nx)          # node columns generated: c0..c1 inclusive
    xs = np.linspace(0., a, nnx)[c0:c1 + 1]
    ys = np.linspace(0., b, nny)
    ncol = xs.size
    X = np.empty((ncol, nny, 3))
    X[:, :, 0] = xs[:, None]
    X[:, :, 1] = ys[None, :]
    X[:, :, 2] = 0.
    X = X.reshape(-1, 3) @ fixed_rotation(seed).T
    ii, jj = np.meshgrid(np.arange(lo, hi), np.arange(ny), indexing="ij")
    n1 = ((ii - c0) * nny + jj).ravel()
    conn = np.stack([n1, n1 + nny, n1 + nny + 1, n1 + 1], 1).astype(np.int64)
    nn_local = ncol * nny
    case = dict(kind=kind, x=X.ravel(), conn=conn, props=shellprop_row(north_star_laminate())[None, :],
                ndof=6 * nn_local, node_offset=c0 * nny)
    if with_u:
        case["u"] = 1e-4 * np.random.default_rng(seed + 7919 * c0).normal(size=6 * nn_local)
    b0 = 0 if i0 is None else i0
    b1 = nnx if i1 is None else i1
    case["owned_nodes"] = ((b0 - c0) * nny, (b1 - c0) * nny)
    case["owned_elements"] = int(((ii >= b0) & (ii < b1)).sum())
    return case


def plate_tria3r(nx, ny, a=0.3, b=0.5, distort=0.4, seed=20):
    """Config 4 (tests/test_tria3r_natural_freq_distorted.py:18-121 in the reference): interior
    nodes perturbed, every quad split into (n1,n2,n3),(n1,n3,n4)."""
    from .shellprop_utils import isotropic_plate
    rng = np.random.default_rng(seed)
    nnx, nny = nx + 1, ny + 1
    xs, ys = np.meshgrid(np.linspace(0, a, nnx), np.linspace(0, b, nny), indexing="ij")
    dx, dy = a / nx, b / ny
    rdm = 2 * rng.random((nnx - 2, nny - 2)) - 1
    xs[1:-1, 1:-1] += distort * dx * rdm
    ys[1:-1, 1:-1] += distort * dy * rdm
    X = np.stack([xs.ravel(), ys.ravel(), np.zeros(nnx * nny)], 1)
    pos = np.arange(nnx * nny).reshape(nnx, nny)
    n1, n2, n3, n4 = pos[:-1, :-1].ravel(), pos[1:, :-1].ravel(), pos[1:, 1:].ravel(), pos[:-1, 1:].ravel()
    conn = np.concatenate([np.stack([n1, n2, n3], 1), np.stack([n1, n3, n4], 1)]).astype(np.int64)
    prop = isotropic_plate(thickness=0.01, E=203e9, nu=0.33, rho=7830.)
    return dict(kind="tria3r", x=X.ravel(), conn=conn, props=shellprop_row(prop)[None, :], ndof=6 * nnx * nny,
                u=1e-5 * rng.normal(size=6 * nnx * nny))


def cylinder_quad4r(ntheta, nlength, L=0.510, R=0.250, seed=0):
    """Config 3 (tests/test_quad4r_linear_buckling_cylinder_Nxy.py:59-115): periodic cylinder,
    laminate, hgfactor 0.001, material axis along the cylinder axis."""
    rng = np.random.default_rng(seed)
    lam = (123.55e9, 8.708e9, 0.319, 5.695e9, 5.695e9, 3.4e9)
    stack = [24, -24, 41, -41]
    prop = laminated_plate(stack=stack, plyt=0.125e-3, laminaprop=lam, rho=1600., calc_scf=False)
    th = np.linspace(0, 2 * np.pi, ntheta, endpoint=False)
    zs = np.linspace(0, L, nlength)
    T, Z = np.meshgrid(th, zs, indexing="ij")
    X = np.stack([R * np.cos(T).ravel(), R * np.sin(T).ravel(), Z.ravel()], 1)
    pos = np.arange(ntheta * nlength).reshape(ntheta, nlength)
    pw = np.vstack([pos, pos[:1]])   # periodic seam
    n1, n2, n3, n4 = pw[:-1, :-1].ravel(), pw[1:, :-1].ravel(), pw[1:, 1:].ravel(), pw[:-1, 1:].ravel()
    conn = np.stack([n1, n2, n3, n4], 1).astype(np.int64)
    ne = conn.shape[0]
    return dict(kind="quad4r", x=X.ravel(), conn=conn, props=shellprop_row(prop)[None, :],
                ndof=6 * ntheta * nlength, xmat=np.tile([0., 0., 1.], (ne, 1)),
                hg=np.full((ne, 5), 0.001), stress=(0., 0., 1000.),
                u=1e-6 * rng.normal(size=6 * ntheta * nlength))


def arc_beamc(n, r=2.438, seed=0):
    """Config 2 (tests/test_beamc_natural_freq_curved.py:20-31): 97-degree arc of n nodes."""
    rng = np.random.default_rng(seed)
    E, G, A, Izz, rho = 206.8e9, 77.9e9 * 5 / 6., 4.071e-3, 6.456e-6, 7855.
    thetas = np.linspace(0, np.radians(97), n)
    X = np.stack([r * np.cos(thetas), r * np.sin(thetas), np.zeros(n)], 1)
    conn = np.stack([np.arange(n - 1), np.arange(1, n)], 1).astype(np.int64)
    p = np.zeros((1, 16))
    Iyy = Izz
    p[0, :9] = [A, E, G, Iyy, Izz, 0., Iyy + Izz, 0., 0.]
    p[0, 9:15] = [rho * A, 0., 0., rho * Izz, rho * Iyy, 0.]
    return dict(kind="beamc", x=X.ravel(), conn=conn, props=p, vxy=np.tile([10., 1., 0.], (n - 1, 1)),
                ndof=6 * n, u=1e-5 * rng.normal(size=6 * n))


def stiffened_panel(nx, ny, nstiff=64, seed=0):
    """Config 5: flat Quad4 skin + BeamC stiffeners along ``nstiff`` grid lines sharing skin nodes."""
    skin = plate_quad4(nx, ny, seed=seed)
    nny = ny + 1
    lines = np.linspace(0, nx, nstiff + 2)[1:-1].round().astype(int)
    i = np.repeat(lines, ny)
    j = np.tile(np.arange(ny), lines.size)
    n1 = i * nny + j
    bconn = np.stack([n1, n1 + 1], 1).astype(np.int64)
    E, nu, rho, bb, hh = 70e9, 0.33, 2700., 0.002, 0.02
    A, Iyy, Izz = bb * hh, bb * hh ** 3 / 12, hh * bb ** 3 / 12
    p = np.zeros((1, 16))
    p[0, :9] = [A, E, E / 2 / (1 + nu) * 5 / 6., Iyy, Izz, 0., Iyy + Izz, 0., 0.]
    p[0, 9:15] = [rho * A, 0., 0., rho * Izz, rho * Iyy, 0.]
    normal = fixed_rotation(seed)[:, 2]
    beams = dict(kind="beamc", x=skin["x"], conn=bconn, props=p, vxy=np.tile(normal, (bconn.shape[0], 1)),
                 ndof=skin["ndof"], u=skin["u"])
    return skin, beams


def batch_from_case(case, device=None):
    """``ElementBatch`` for a case dict returned by the generators above."""
    from .batch import ElementBatch
    kind = case["kind"]
    kw = dict(x=case.get("x"), props=case.get("props"), prop_id=case.get("prop_id"), u=case.get("u"),
              nnodes=case["ndof"] // 6, device=device)
    if kind in ("quad4", "quad4r", "tria3r"):
        kw.update(xmat=case.get("xmat"), K6ROT=case.get("K6ROT"), alpha_shear_locking=case.get("alpha"),
                  hgfactors=case.get("hg"))
    elif kind in ("beamc", "beamlr"):
        kw.update(vxy=case["vxy"])
    elif kind == "spring":
        kw.update(axes=case["axes"], k=case["k"])
    return ElementBatch(kind, case["conn"], **kw)


def static_case(side):
    """The bounded static-solve workload both arms of bench.py's ``config.solve_e2e`` run: the workload mesh at
    side x side elements, all six dofs clamped on the four edges, a unit load along the plate normal on every
    interior node (the flow of tests/test_quad4_static_point_load.py:53-104 with the diagonal scaling of
    tests/test_quad4r_linear_buckling_plate.py:135-146)."""
    case = plate_quad4(side, side, with_u=False)
    nny = side + 1
    nn = case["ndof"] // 6
    i, j = np.divmod(np.arange(nn), nny)
    edge = (i == 0) | (i == side) | (j == 0) | (j == side)
    free = np.repeat(~edge, 6)
    normal = fixed_rotation(0)[:, 2]
    f = np.zeros(case["ndof"])
    for d in range(3):
        f[d::6] = normal[d] * (~edge)
    return case, free, f, normal
