"""Seeded synthetic cases shared by the golden generator, the oracle tests and
the GPU parity tests.  A *case* is the plain dict documented in
``oracle/driver.py``."""
import numpy as np


def random_rotation(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    return q


def random_shellprops(rng, n, coupled=True):
    """Random symmetric-positive-definite ABD (B != 0) + shear + inertia."""
    P = np.zeros((n, 32))
    for i in range(n):
        L = rng.normal(size=(6, 6))
        C = L @ L.T + 6 * np.eye(6)
        h = 10 ** rng.uniform(-3, -1.5)
        sA, sD = 70e9 * h, 70e9 * h ** 3 / 12
        s = np.sqrt(np.array([sA] * 3 + [sD] * 3))
        C = C * s[:, None] * s[None, :] / 6.
        if not coupled:
            C[:3, 3:] = 0
            C[3:, :3] = 0
        A, B, D = C[:3, :3], C[:3, 3:], C[3:, 3:]
        B = 0.5 * (B + B.T)
        iu = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
        P[i, 0:6] = [A[a] for a in iu]
        P[i, 6:12] = [B[a] for a in iu]
        P[i, 12:18] = [D[a] for a in iu]
        P[i, 18:21] = 27e9 * h * np.array([1.0, 0.05, 1.1])
        P[i, 21:23] = [5 / 6., 0.8]
        P[i, 23] = h
        rho = 2700.
        P[i, 24:27] = [rho * h, rho * h * h * 0.1, rho * h ** 3 / 12]
    return P


def random_beamprops(rng, n, full=True):
    P = np.zeros((n, 16))
    for i in range(n):
        b, hh = rng.uniform(0.01, 0.05, 2)
        A = b * hh
        Iyy, Izz = b * hh ** 3 / 12, hh * b ** 3 / 12
        E = 70e9
        nu = 0.3
        G = E / 2 / (1 + nu)
        rho = 2700.
        Iyz = 0.2 * np.sqrt(Iyy * Izz) if full else 0.
        Ay, Az = (0.1 * A * b, -0.07 * A * hh) if full else (0., 0.)
        P[i, :9] = [A, E, G, Iyy, Izz, Iyz, Iyy + Izz, Ay, Az]
        P[i, 9:15] = [rho * A, rho * Ay, rho * Az, rho * Izz, rho * Iyy, rho * Iyz]
    return P


def _finish(case, nnodes, rng, uscale=1e-4):
    case["ndof"] = 6 * nnodes
    case["u"] = uscale * rng.normal(size=6 * nnodes)
    return case


def shell_soup(kind, ne, seed, size_range=(-2.5, 0.), thick=False, xmat=True, nprop=3,
               offset_scale=2.0):
    """ne disconnected, randomly rotated/distorted/translated elements.  The
    translation is ``offset_scale`` element sizes: the reference forms local
    coordinates from ABSOLUTE positions (quad4.pyx:724-728), so elements far from
    the origin relative to their size lose digits in *both* implementations
    (SURVEY §7 'hard parts'); parity cases keep that conditioning benign."""
    rng = np.random.default_rng(seed)
    nn = 3 if kind == "tria3r" else 4
    base = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], float)[:nn]
    if nn == 3:
        base = np.array([[0, 0, 0], [1, 0, 0], [0.3, 0.9, 0]], float)
    X = np.zeros((ne * nn, 3))
    for e in range(ne):
        s = 10 ** rng.uniform(*size_range)
        pts = base + 0.18 * rng.uniform(-1, 1, (nn, 3)) * [1, 1, 0.1]
        pts = s * pts * [1., rng.uniform(0.5, 2.), 1.]
        X[e * nn:(e + 1) * nn] = pts @ random_rotation(rng).T + offset_scale * s * rng.normal(size=3)
    conn = np.arange(ne * nn, dtype=np.int64).reshape(ne, nn)
    props = random_shellprops(rng, nprop)
    if thick:
        props[:, 23] = 10 ** size_range[1] * 3.
    case = dict(kind=kind, x=X.ravel(), conn=conn, props=props,
                prop_id=rng.integers(0, nprop, ne).astype(np.int32),
                K6ROT=rng.choice([100., 1e4], ne), stress=(1.3e3, -0.7e3, 0.4e3))
    if kind == "tria3r":
        case["alpha"] = rng.uniform(0.3, 1.0, ne)
    if kind == "quad4r":
        case["hg"] = rng.uniform(0.001, 2.0, (ne, 5))
    if xmat:
        xm = rng.normal(size=(ne, 3))
        xm[::5] = 0.                      # no material axis -> identity m
        case["xmat"] = xm
    return _finish(case, ne * nn, rng)


def shell_mesh(kind, nx, ny, seed, a=1.3, b=0.8, distort=0.25, curved=True):
    """Connected structured mesh (shared nodes), rigidly rotated."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.linspace(0, a, nx), np.linspace(0, b, ny), indexing="ij")
    dx, dy = a / (nx - 1), b / (ny - 1)
    xs = xs + distort * dx * rng.uniform(-1, 1, xs.shape)
    ys = ys + distort * dy * rng.uniform(-1, 1, ys.shape)
    zs = 0.05 * np.sin(3 * xs) * np.cos(2 * ys) if curved else np.zeros_like(xs)
    X = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], 1) @ random_rotation(rng).T + [0.3, -0.2, 0.1]
    pos = np.arange(nx * ny).reshape(nx, ny)
    n1, n2, n3, n4 = pos[:-1, :-1].ravel(), pos[1:, :-1].ravel(), pos[1:, 1:].ravel(), pos[:-1, 1:].ravel()
    if kind == "tria3r":
        conn = np.concatenate([np.stack([n1, n2, n3], 1), np.stack([n1, n3, n4], 1)])
    else:
        conn = np.stack([n1, n2, n3, n4], 1)
    conn = conn.astype(np.int64)
    ne = conn.shape[0]
    props = random_shellprops(rng, 2)
    case = dict(kind=kind, x=X.ravel(), conn=conn, props=props,
                prop_id=rng.integers(0, 2, ne).astype(np.int32), stress=(-1e3, 0., 2e2))
    case["xmat"] = np.tile(rng.normal(size=3), (ne, 1))
    return _finish(case, nx * ny, rng)


def line_soup(kind, ne, seed, full=True, logL=(-1.5, 0.5)):
    """Disconnected random line elements.  BeamC divides by (1 - alpha), alpha = 12EI/(GAL^2)
    (beamc.pyx:543-546): stubby random beams with alpha ~ 1 are ill-conditioned in the reference
    itself, so large random sweeps use a longer ``logL`` range."""
    rng = np.random.default_rng(seed)
    X = np.zeros((2 * ne, 3))
    vxy = np.zeros((ne, 3))
    for e in range(ne):
        L = 10 ** rng.uniform(*logL)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        p0 = rng.normal(size=3)
        X[2 * e], X[2 * e + 1] = p0, p0 + L * d
        v = rng.normal(size=3)
        vxy[e] = v
    conn = np.arange(2 * ne, dtype=np.int64).reshape(ne, 2)
    case = dict(kind=kind, x=X.ravel(), conn=conn)
    if kind == "spring":
        case["k"] = 10 ** rng.uniform(3, 7, (ne, 6))
        case["axes"] = np.concatenate([rng.normal(size=(ne, 3)), rng.normal(size=(ne, 3))], 1)
        case["props"] = None
    else:
        case["props"] = random_beamprops(rng, 3, full=full)
        case["prop_id"] = rng.integers(0, 3, ne).astype(np.int32)
        case["vxy"] = vxy
    return _finish(case, 2 * ne, rng)


def line_chain(kind, n, seed):
    """Connected curved chain of n-1 line elements (shared nodes)."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0, 1.7, n)
    X = np.stack([2.4 * np.cos(t), 2.4 * np.sin(t), 0.3 * t], 1) @ random_rotation(rng).T
    conn = np.stack([np.arange(n - 1), np.arange(1, n)], 1).astype(np.int64)
    ne = n - 1
    case = dict(kind=kind, x=X.ravel(), conn=conn)
    if kind == "spring":
        case["k"] = 10 ** rng.uniform(3, 7, (ne, 6))
        case["axes"] = np.concatenate([rng.normal(size=(ne, 3)), rng.normal(size=(ne, 3))], 1)
        case["props"] = None
    else:
        case["props"] = random_beamprops(rng, 1)
        case["vxy"] = np.tile(rng.normal(size=3), (ne, 1))
    return _finish(case, n, rng)


def shell_degenerate_xmat(kind, seed):
    """The guards of update_rotation_matrix (quad4.pyx:588-598, tria3r.pyx:388-398; the situation of
    tests/test_quad4r_static_point_load_mat_coord_error.py:74 in the reference): material direction EXACTLY parallel
    to the element normal (flat elements in the global xy plane: z = (0,0,1) without rounding), parallel up to
    rounding (xmat = the computed normal of a rotated element, and the exact normal plus a perturbation far below
    tol = |z|/1e10), antiparallel, null, and -- as controls -- in-plane and generic directions.  In every degenerate
    row the reference leaves m at its previous value, which is the identity for a fresh element."""
    rng = np.random.default_rng(seed)
    nn = 3 if kind == "tria3r" else 4
    base = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], float)[:nn]
    if nn == 3:
        base = np.array([[0, 0, 0], [1, 0, 0], [0.3, 0.9, 0]], float)
    ne = 14
    X = np.zeros((ne * nn, 3))
    xm = np.zeros((ne, 3))
    for e in range(ne):
        s = 10 ** rng.uniform(-1.5, 0.)
        pts = s * (base + 0.15 * rng.uniform(-1, 1, (nn, 3)) * [1, 1, 0]) + s * rng.normal(size=3) * [1, 1, 0]
        if e < 8:
            X[e * nn:(e + 1) * nn] = pts                      # flat, unrotated: normal = +z exactly
        else:
            X[e * nn:(e + 1) * nn] = pts @ random_rotation(rng).T
    conn = np.arange(ne * nn, dtype=np.int64).reshape(ne, nn)
    P = X.reshape(ne, nn, 3)
    if nn == 4:
        nrm = np.cross(P[:, 1] - P[:, 3], P[:, 2] - P[:, 0])
    else:
        nrm = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
    unit = nrm / np.linalg.norm(nrm, axis=1)[:, None]
    xm[0] = [0., 0., 1.]
    xm[1] = [0., 0., 2.5]
    xm[2] = [0., 0., -3.]
    xm[3] = [0., 0., 0.]
    xm[4] = [1e-13, -2e-13, 1.]          # |z x xmat| ~ 2e-13 < tol
    xm[5] = [1., 0., 0.]                 # in plane: m = identity through the regular branch
    xm[6] = [0.3, 0.7, 0.]
    xm[7] = [0.3, 0.7, 5.]
    xm[8] = unit[8]                      # computed normals of rotated elements: parallel up to rounding
    xm[9] = -4. * unit[9]
    xm[10] = nrm[10]                     # unnormalised
    xm[11] = unit[11] + 1e-14 * rng.normal(size=3)
    xm[12] = rng.normal(size=3)          # generic controls
    xm[13] = rng.normal(size=3)
    props = random_shellprops(rng, 2)
    case = dict(kind=kind, x=X.ravel(), conn=conn, props=props, prop_id=rng.integers(0, 2, ne).astype(np.int32),
                stress=(0.9e3, 0.2e3, -0.5e3), xmat=xm)
    if kind == "quad4r":
        case["hg"] = rng.uniform(0.001, 2.0, (ne, 5))
    return _finish(case, ne * nn, rng)


SHELL_KINDS = ("quad4", "quad4r", "tria3r")
LINE_KINDS = ("beamc", "beamlr", "truss", "spring")


def golden_cases():
    """name -> case; everything the committed fixtures cover."""
    out = {}
    for k in SHELL_KINDS:
        out[k + "_soup"] = shell_soup(k, 24, seed=11)
        out[k + "_soup_thick"] = shell_soup(k, 8, seed=12, size_range=(-3., -2.5), thick=True)
        out[k + "_mesh"] = shell_mesh(k, 5, 4, seed=13)
        out[k + "_xmat_degenerate"] = shell_degenerate_xmat(k, seed=14)
    for k in LINE_KINDS:
        out[k + "_soup"] = line_soup(k, 24, seed=21)
        out[k + "_chain"] = line_chain(k, 9, seed=22)
    return out
