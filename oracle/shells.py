"""numpy restatement of the reference shell elements (Quad4, Quad4R, Tria3R).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Everything is vectorised over a batch of ``ne`` elements and written in the
textbook B-matrix form (``Ke = sum_gp detJ * B^T C B``), deliberately *not* the
factored form the CUDA kernels use, so the two are independent statements of
the same mathematics.  Citations are to ``/root/reference/pyfe3d``.

Array conventions (reference: quad4.pyx:526-537, 634-639):
  x     float64[3*nnodes]   global coordinates, node p at x[3p:3p+3]
  u     float64[6*nnodes]   global DOFs (u v w rx ry rz per node)
  conn  int64[ne, nn]       node *positions* p; the reference's c_a = 6*p
  props float64[nprop, 32]  ShellProp scalars, layout SHELLPROP_FIELDS
"""
import numpy as np

SHELLPROP_FIELDS = ["A11", "A12", "A16", "A22", "A26", "A66",
                    "B11", "B12", "B16", "B22", "B26", "B66",
                    "D11", "D12", "D16", "D22", "D26", "D66",
                    "E44", "E45", "E55", "scf_k13", "scf_k23", "h",
                    "intrho", "intrhoz", "intrhoz2"]
SHELLPROP_STRIDE = 32
GP = 0.5773502691896257645092  # quad4.pyx:929-930


def pack_shellprops(objs):
    """ShellProp-like objects (attribute access) -> float64[nprop, 32]."""
    out = np.zeros((len(objs), SHELLPROP_STRIDE))
    for i, o in enumerate(objs):
        for j, f in enumerate(SHELLPROP_FIELDS):
            out[i, j] = getattr(o, f)
    return out


# --------------------------------------------------------------------------
# frames
# --------------------------------------------------------------------------
def _unit(v):
    n = np.sqrt((v * v).sum(-1))
    return v / n[:, None], n


def _material_axes(R, znorm, xmat, m_prev):
    """quad4.pyx:583-624 / tria3r.pyx same block.  m is sticky: untouched when
    xmat is null or (nearly) parallel to the normal."""
    ne = R.shape[0]
    m = np.tile(np.eye(2), (ne, 1, 1)) if m_prev is None else m_prev.copy()
    if xmat is None:
        return m
    xmat = np.broadcast_to(np.asarray(xmat, float), (ne, 3)).copy()
    xh, yh, zh = R[:, :, 0], R[:, :, 1], R[:, :, 2]
    tol = znorm / 1e10
    with np.errstate(all="ignore"):
        xn = np.sqrt((xmat ** 2).sum(1))
        xm = xmat / xn[:, None]
        ymat = np.cross(zh, xm)
        yn = np.sqrt((ymat ** 2).sum(1))
        ymat = ymat / yn[:, None]
        xp = np.cross(ymat, zh)
        xp = xp / np.sqrt((xp ** 2).sum(1))[:, None]
        cost = (xp * xh).sum(1)
        sint = np.sqrt(1 - cost ** 2)
        pos = (xp * yh).sum(1) > 0
    ok = (xn > tol) & (yn > tol)
    m12 = np.where(pos, -sint, sint)
    for e in np.nonzero(ok)[0]:
        m[e] = [[cost[e], m12[e]], [-m12[e], cost[e]]]
    return m


def quad_frames(x, conn, xmat=None, m_prev=None):
    """Quad4/Quad4R.update_rotation_matrix (quad4.pyx:491-624, quad4r.pyx:286-419).
    Returns R[ne,3,3] (columns = element x,y,z in global) and m[ne,2,2]."""
    X = x.reshape(-1, 3)[conn]
    v13 = X[:, 2] - X[:, 0]
    v42 = X[:, 1] - X[:, 3]
    z, zn = _unit(np.cross(v42, v13))
    xh, _ = _unit((v13 + v42) / 2.)
    yh, _ = _unit(np.cross(z, xh))
    R = np.stack([xh, yh, z], axis=2)
    return R, _material_axes(R, zn, xmat, m_prev)


def tria_frames(x, conn, xmat=None, m_prev=None):
    """Tria3R.update_rotation_matrix (tria3r.pyx:294-424)."""
    X = x.reshape(-1, 3)[conn]
    v12 = X[:, 1] - X[:, 0]
    v13 = X[:, 2] - X[:, 0]
    z, zn = _unit(np.cross(v12, v13))
    xh, _ = _unit(v12)
    yh, _ = _unit(np.cross(z, xh))
    R = np.stack([xh, yh, z], axis=2)
    return R, _material_axes(R, zn, xmat, m_prev)


def local_xe(R, x, conn):
    """update_probe_xe: xe_a = R^T x_a on ABSOLUTE coordinates (quad4.pyx:682-728)."""
    X = x.reshape(-1, 3)[conn]
    return np.einsum("eji,eaj->eai", R, X)


def local_ue(R, u, conn):
    """update_probe_ue (quad4.pyx:627-679): ue = (R^T u_trans, R^T u_rot) per node."""
    U = u.reshape(-1, 2, 3)[conn]                       # e a (t|r) j
    return np.einsum("eji,eatj->eati", R, U).reshape(conn.shape[0], -1)


def quad_area(xe):
    """quad4.pyx:733-752 (shoelace on local x,y)."""
    X, Y = xe[:, :, 0], xe[:, :, 1]
    return 0.5 * np.abs((X[:, 0] * Y[:, 1] + X[:, 1] * Y[:, 2] + X[:, 2] * Y[:, 3] + X[:, 3] * Y[:, 0])
                        - (X[:, 1] * Y[:, 0] + X[:, 2] * Y[:, 1] + X[:, 3] * Y[:, 2] + X[:, 0] * Y[:, 3]))


def tria_area(xe):
    """tria3r.pyx:530-546."""
    X, Y = xe[:, :, 0], xe[:, :, 1]
    return np.abs((-X[:, 0] + X[:, 1]) * (-Y[:, 0] + Y[:, 2]) / 2. + (X[:, 0] - X[:, 2]) * (-Y[:, 0] + Y[:, 1]) / 2.)


# --------------------------------------------------------------------------
# constitutive data
# --------------------------------------------------------------------------
def _sym3(p6):
    a11, a12, a16, a22, a26, a66 = [p6[:, i] for i in range(6)]
    return np.stack([np.stack([a11, a12, a16], 1), np.stack([a12, a22, a26], 1),
                     np.stack([a16, a26, a66], 1)], 1)


def abd_element_axes(pe, m):
    """A,B,D rotated from material to element axes when m12 != 0
    (quad4.pyx:847-899).  E44/45/55 are never rotated."""
    A, B, D = _sym3(pe[:, 0:6]), _sym3(pe[:, 6:12]), _sym3(pe[:, 12:18])
    m11, m12, m21, m22 = m[:, 0, 0], m[:, 0, 1], m[:, 1, 0], m[:, 1, 1]
    Tm = np.stack([np.stack([m11 ** 2, m12 ** 2, 2 * m11 * m12], 1),
                   np.stack([m21 ** 2, m22 ** 2, 2 * m21 * m22], 1),
                   np.stack([m11 * m21, m12 * m22, m11 * m22 + m12 * m21], 1)], 1)
    rot = (m12 != 0)[:, None, None]
    out = []
    for M in (A, B, D):
        out.append(np.where(rot, Tm @ M @ Tm.transpose(0, 2, 1), M))
    return out


def shear_moduli(pe):
    """quad4.pyx:903-905."""
    E44 = pe[:, 18] * pe[:, 22]
    E45 = pe[:, 19] * 0.5 * (pe[:, 21] + pe[:, 22])
    E55 = pe[:, 20] * pe[:, 21]
    return np.stack([np.stack([E44, E45], 1), np.stack([E45, E55], 1)], 1)


# --------------------------------------------------------------------------
# quad shape data
# --------------------------------------------------------------------------
_XI = np.array([-1., 1., 1., -1.])
_ETA = np.array([-1., -1., 1., 1.])


def quad_shape(xe, xi, eta):
    """N[4], Nx[ne,4], Ny[ne,4], detJ[ne], jinv[ne,2,2] at (xi, eta)
    (Quad4Probe.update_BL, quad4.pyx:299-322)."""
    N = 0.25 * (1 + _XI * xi) * (1 + _ETA * eta)
    dxi = 0.25 * _XI * (1 + _ETA * eta)
    deta = 0.25 * _ETA * (1 + _XI * xi)
    X, Y = xe[:, :, 0], xe[:, :, 1]
    J11, J12 = X @ dxi, Y @ dxi
    J21, J22 = X @ deta, Y @ deta
    det = J11 * J22 - J12 * J21
    j11, j12, j21, j22 = J22 / det, -J12 / det, -J21 / det, J11 / det
    Nx = j11[:, None] * dxi + j12[:, None] * deta
    Ny = j21[:, None] * dxi + j22[:, None] * deta
    jinv = np.stack([np.stack([j11, j12], 1), np.stack([j21, j22], 1)], 1)
    return N, Nx, Ny, det, jinv


def strain_rows(N, Nx, Ny, nn):
    """B rows per SURVEY App. A / quad4.pyx:324-395.  N may be [nn] or [ne,nn]."""
    ne = Nx.shape[0]
    N = np.broadcast_to(N, (ne, nn))
    B = np.zeros((ne, 6, 6 * nn))
    Bs = np.zeros((ne, 2, 6 * nn))
    Bg = np.zeros((ne, 2, 6 * nn))
    Bd = np.zeros((ne, 6 * nn))
    for a in range(nn):
        o = 6 * a
        B[:, 0, o + 0] = Nx[:, a]
        B[:, 1, o + 1] = Ny[:, a]
        B[:, 2, o + 0] = Ny[:, a]
        B[:, 2, o + 1] = Nx[:, a]
        B[:, 3, o + 4] = Nx[:, a]
        B[:, 4, o + 3] = -Ny[:, a]
        B[:, 5, o + 3] = -Nx[:, a]
        B[:, 5, o + 4] = Ny[:, a]
        Bs[:, 0, o + 2] = Ny[:, a]
        Bs[:, 0, o + 3] = -N[:, a]
        Bs[:, 1, o + 2] = Nx[:, a]
        Bs[:, 1, o + 4] = N[:, a]
        Bg[:, 0, o + 2] = Ny[:, a]
        Bg[:, 1, o + 2] = Nx[:, a]
        Bd[:, o + 0] = Ny[:, a] / 2.
        Bd[:, o + 1] = -Nx[:, a] / 2.
        Bd[:, o + 5] = N[:, a]
    return B, Bs, Bg, Bd


def _C6(A, B, D):
    return np.block([[A, B], [B, D]])


def _btcb(B, C):
    return np.einsum("eki,ekl,elj->eij", B, C, B)


# --------------------------------------------------------------------------
# local stiffness matrices
# --------------------------------------------------------------------------
def quad4_Ke(xe, area, pe, m):
    """Quad4._update_probe_KC0ve (quad4.pyx:755-1171).  Drilling coefficient is
    1.0 (K6ROT is ignored by the reference)."""
    A, B_, D = abd_element_axes(pe, m)
    C = _C6(A, B_, D)
    E = shear_moduli(pe)
    ne = xe.shape[0]
    thick = (pe[:, 23] / np.sqrt(area) >= 1.)[:, None, None]
    Ke = np.zeros((ne, 24, 24))
    for xi in (-GP, GP):
        for eta in (-GP, GP):
            N, Nx, Ny, det, _ = quad_shape(xe, xi, eta)
            B, Bs, Bg, Bd = strain_rows(N, Nx, Ny, 4)
            Kgp = _btcb(B, C) + np.einsum("ei,ej->eij", Bd, Bd)
            Kgp = Kgp + np.where(thick, _btcb(Bg, E), 0.)
            Ke += det[:, None, None] * Kgp
    N, Nx, Ny, det, _ = quad_shape(xe, 0., 0.)
    B, Bs, Bg, Bd = strain_rows(N, Nx, Ny, 4)
    Kc = _btcb(Bs, E) - np.where(thick, _btcb(Bg, E), 0.)
    Ke += 4. * det[:, None, None] * Kc
    return Ke


def quad4r_Ke(xe, area, pe, m, K6ROT, hg):
    """Quad4R.update_KC0 local matrix (quad4r.pyx:1253-3464).  hg[ne,5] =
    hgfactor_u,v,w,rx,ry."""
    A, B_, D = abd_element_axes(pe, m)
    C = _C6(A, B_, D)
    E = shear_moduli(pe)
    h = pe[:, 23]
    Ainv = np.linalg.inv(A)
    E1eq = 1. / (h * Ainv[:, 0, 0])
    E2eq = 1. / (h * Ainv[:, 1, 1])
    den = 1.0 + 1.0 / area
    Eu = hg[:, 0] * 0.1 * E1eq * h / den
    Ev = hg[:, 1] * 0.1 * E2eq * h / den
    Erx = hg[:, 3] * 0.1 * E2eq * h ** 3 / den
    Ery = hg[:, 4] * 0.1 * E1eq * h ** 3 / den
    Ew = hg[:, 2] * 0.5 * (Erx + Ery)
    N, Nx, Ny, det, j = quad_shape(xe, 0., 0.)
    B, Bs, Bg, Bd = strain_rows(N, Nx, Ny, 4)
    g = 0.25 * (j[:, 0, 0] * j[:, 1, 1] + j[:, 0, 1] * j[:, 1, 0])
    gam = g[:, None] * np.array([1., -1., 1., -1.])
    ne = xe.shape[0]
    H = np.zeros((ne, 24, 24))
    for d, Ed in enumerate((Eu, Ev, Ew, Erx, Ery)):
        for a in range(4):
            for b in range(4):
                H[:, 6 * a + d, 6 * b + d] = Ed * gam[:, a] * gam[:, b]
    Ke = 4. * det[:, None, None] * (_btcb(B, C) + _btcb(Bs, E) + H)
    A66 = A[:, 2, 2]
    for xi in (-GP, GP):
        for eta in (-GP, GP):
            N, Nx, Ny, det, _ = quad_shape(xe, xi, eta)
            _, _, _, Bd = strain_rows(N, Nx, Ny, 4)
            Ke += (det * 1e-6 * K6ROT * A66)[:, None, None] * np.einsum("ei,ej->eij", Bd, Bd)
    return Ke


_TRIA_PTS = np.array([[2 / 3., 1 / 6., 1 / 6.], [1 / 6., 1 / 6., 2 / 3.], [1 / 6., 2 / 3., 1 / 6.]])


def tria_grads(xe, area):
    """tria3r.pyx:2116-2121."""
    X, Y = xe[:, :, 0], xe[:, :, 1]
    a2 = 2 * area
    Nx = np.stack([Y[:, 1] - Y[:, 2], -Y[:, 0] + Y[:, 2], Y[:, 0] - Y[:, 1]], 1) / a2[:, None]
    Ny = np.stack([-X[:, 1] + X[:, 2], X[:, 0] - X[:, 2], -X[:, 0] + X[:, 1]], 1) / a2[:, None]
    return Nx, Ny


def tria3r_Ke(xe, area, pe, m, K6ROT, alpha, drop_drilling_couplings):
    """Tria3R local stiffness (tria3r.pyx:950-2325).  ``drop_drilling_couplings``
    reproduces update_KC0's quirk (:2325-2973 never reads the (u,v)-rz and
    rz_a-rz_b drilling scalars); update_probe_finte (:549-948) keeps them."""
    A, B_, D = abd_element_axes(pe, m)
    C = _C6(A, B_, D)
    E = shear_moduli(pe)
    h = pe[:, 23]
    X, Y = xe[:, :, 0], xe[:, :, 1]
    l12 = np.sqrt((X[:, 0] - X[:, 1]) ** 2 + (Y[:, 0] - Y[:, 1]) ** 2)
    l23 = np.sqrt((X[:, 1] - X[:, 2]) ** 2 + (Y[:, 1] - Y[:, 2]) ** 2)
    l31 = np.sqrt((X[:, 2] - X[:, 0]) ** 2 + (Y[:, 2] - Y[:, 0]) ** 2)
    maxl = np.maximum(np.maximum(l12, l23), l31)
    factor = alpha * maxl ** 2 / h ** 2
    E = E * (1. / (1. + factor))[:, None, None]
    detJ = 2 * area
    Nx, Ny = tria_grads(xe, area)
    B, Bs, _, _ = strain_rows(np.full(3, 1 / 3.), Nx, Ny, 3)
    Ke = (detJ * 0.5)[:, None, None] * (_btcb(B, C) + _btcb(Bs, E))
    A66 = A[:, 2, 2]
    mask = np.ones((18, 18))
    if drop_drilling_couplings:
        rz = np.zeros(18, bool)
        rz[[5, 11, 17]] = True
        mask[np.ix_(rz, ~rz)] = 0
        mask[np.ix_(~rz, rz)] = 0
        mask[np.ix_(rz, rz)] = np.eye(3)
    for p in range(3):
        _, _, _, Bd = strain_rows(_TRIA_PTS[p], Nx, Ny, 3)
        Ke += (detJ * (0.5 / 3.) * 1e-6 * K6ROT * A66)[:, None, None] * mask * np.einsum("ei,ej->eij", Bd, Bd)
    return Ke


# --------------------------------------------------------------------------
# geometric stiffness and mass (local)
# --------------------------------------------------------------------------
def _Ge_to_local(Ge, nn):
    Kl = np.zeros((Ge.shape[0], 6 * nn, 6 * nn))
    Kl[:, 2::6, 2::6] = Ge
    return Kl


def quad_KG_local(xe, pe, m, ue=None, stress=None):
    """Quad4/Quad4R update_KG (quad4.pyx:1365-2257) and update_KG_given_stress
    (:2259-3081).  Acts on w only; 2x2 Gauss."""
    ne = xe.shape[0]
    Ge = np.zeros((ne, 4, 4))
    if stress is None:
        A, B_, D = abd_element_axes(pe, m)
    for xi in (-GP, GP):
        for eta in (-GP, GP):
            N, Nx, Ny, det, _ = quad_shape(xe, xi, eta)
            if stress is None:
                B, _, _, _ = strain_rows(N, Nx, Ny, 4)
                eps = np.einsum("eki,ei->ek", B, ue)
                Nm = np.einsum("eij,ej->ei", A, eps[:, :3]) + np.einsum("eij,ej->ei", B_, eps[:, 3:])
                Nxx, Nyy, Nxy = Nm[:, 0], Nm[:, 1], Nm[:, 2]
            else:
                Nxx, Nyy, Nxy = [np.broadcast_to(s, (ne,)) for s in stress]
            Ge += det[:, None, None] * (
                Nx[:, :, None] * (Nx[:, None, :] * Nxx[:, None, None] + Ny[:, None, :] * Nxy[:, None, None])
                + Ny[:, :, None] * (Nx[:, None, :] * Nxy[:, None, None] + Ny[:, None, :] * Nyy[:, None, None]))
    return _Ge_to_local(Ge, 4)


def quad_aero_local(xe, R):
    """Piston-theory aerodynamic matrices of Quad4 / Quad4R: update_KA_beta (quad4.pyx:9491-10310,
    quad4r.pyx:12789-13603), update_KA_gamma (quad4.pyx:10312-11113, quad4r.pyx:13605-14401) and update_CA
    (quad4.pyx:11115-11917, quad4r.pyx:14403-15215).  Like KG they act on w only, 2x2 Gauss, wij = 1:
        KA_beta [w_a, w_b] = - sum_gp N_a detJ (N_b,x r11 + N_b,y r21)      (quad4.pyx:9687 ff.)
        KA_gamma[w_a, w_b] =   sum_gp N_a N_b detJ                           (quad4.pyx:10498 ff.)
        CA      [w_a, w_b] = - sum_gp N_a N_b detJ                           (quad4.pyx:11301 ff.)
    and the global 3x3 block (a, b) is that scalar times r_{.3} r_{.3}^T.  Returns three local 24x24 matrices."""
    ne = xe.shape[0]
    Hb = np.zeros((ne, 4, 4))
    Hg = np.zeros((ne, 4, 4))
    r11, r21 = R[:, 0, 0], R[:, 1, 0]
    for xi in (-GP, GP):
        for eta in (-GP, GP):
            N, Nx, Ny, det, _ = quad_shape(xe, xi, eta)
            flow = Nx * r11[:, None] + Ny * r21[:, None]
            Hb -= det[:, None, None] * N[None, :, None] * flow[:, None, :]
            Hg += det[:, None, None] * np.outer(N, N)
    return _Ge_to_local(Hb, 4), _Ge_to_local(Hg, 4), _Ge_to_local(-Hg, 4)


def tria_KG_local(xe, area, pe, m, ue=None, stress=None):
    """Tria3R update_KG (tria3r.pyx:3020-3574) / given stress (:3576-4061); 1 point, w=1/2."""
    ne = xe.shape[0]
    Nx, Ny = tria_grads(xe, area)
    if stress is None:
        A, B_, D = abd_element_axes(pe, m)
        B, _, _, _ = strain_rows(np.full(3, 1 / 3.), Nx, Ny, 3)
        eps = np.einsum("eki,ei->ek", B, ue)
        Nm = np.einsum("eij,ej->ei", A, eps[:, :3]) + np.einsum("eij,ej->ei", B_, eps[:, 3:])
        Nxx, Nyy, Nxy = Nm[:, 0], Nm[:, 1], Nm[:, 2]
    else:
        Nxx, Nyy, Nxy = [np.broadcast_to(s, (ne,)) for s in stress]
    Ge = (2 * area * 0.5)[:, None, None] * (
        Nx[:, :, None] * (Nx[:, None, :] * Nxx[:, None, None] + Ny[:, None, :] * Nxy[:, None, None])
        + Ny[:, :, None] * (Nx[:, None, :] * Nxy[:, None, None] + Ny[:, None, :] * Nyy[:, None, None]))
    return _Ge_to_local(Ge, 3)


def _ml(pe, lumped):
    """6x6 nodal inertia block (quad4.pyx:4644 ff.; SURVEY §8(a) update_M row)."""
    r0, r1, r2 = pe[:, 24], pe[:, 25], pe[:, 26]
    ml = np.zeros((pe.shape[0], 6, 6))
    for d in range(3):
        ml[:, d, d] = r0
    ml[:, 3, 3] = r2
    ml[:, 4, 4] = r2
    if not lumped:
        ml[:, 0, 4] = ml[:, 4, 0] = r1
        ml[:, 1, 3] = ml[:, 3, 1] = -r1
    return ml


def quad_M_local(xe, area, pe, mtype):
    """Quad4/Quad4R.update_M (quad4.pyx:3083-9489)."""
    ne = xe.shape[0]
    H = np.zeros((ne, 4, 4))
    if mtype == 0:
        for xi in (-GP, GP):
            for eta in (-GP, GP):
                N, _, _, det, _ = quad_shape(xe, xi, eta)
                H += det[:, None, None] * np.outer(N, N)
    elif mtype == 1:
        H += (0.0625 * area)[:, None, None]
    else:
        for xi in (-1., 1.):
            for eta in (-1., 1.):
                N, _, _, det, _ = quad_shape(xe, xi, eta)
                H += det[:, None, None] * np.outer(N, N)
    return np.einsum("eab,eij->eaibj", H, _ml(pe, mtype == 2)).reshape(ne, 24, 24)


def tria_M_local(xe, area, pe, mtype):
    """Tria3R.update_M (tria3r.pyx:4063-7723)."""
    ne = xe.shape[0]
    detJ = 2 * area
    H = np.zeros((ne, 3, 3))
    if mtype == 0:
        for p in range(3):
            H += (detJ * 0.5 / 3.)[:, None, None] * np.outer(_TRIA_PTS[p], _TRIA_PTS[p])
    elif mtype == 1:
        H += (detJ / 18.)[:, None, None]
    else:
        H += (detJ * 0.5 / 3.)[:, None, None] * np.eye(3)
    return np.einsum("eab,eij->eaibj", H, _ml(pe, mtype == 2)).reshape(ne, 18, 18)
