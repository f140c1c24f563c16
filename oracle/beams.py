"""numpy restatement of the reference line elements (BeamC, BeamLR, Truss, Spring).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Local 12x12 matrices are stated as sparse upper-triangular term lists (DOF order
u1 v1 w1 rx1 ry1 rz1 u2 ...), symmetrised, then rotated with T = blockdiag(R x4).
Sources: beamc.pyx:543-624 (KC0e), :1889-2179 (KGe), :2246-3152 (Me);
beamlr.pyx:460-1186, :1388-1462, :1518-2424; truss.pyx:434-800, :894-1800;
spring.pyx:340-706.  SURVEY Appendix A lists the recovered tables.
"""
import numpy as np

from . import coo

BEAMPROP_FIELDS = ["A", "E", "G", "Iyy", "Izz", "Iyz", "J", "Ay", "Az",
                   "intrho", "intrhoy", "intrhoz", "intrhoy2", "intrhoz2", "intrhoyz"]
BEAMPROP_STRIDE = 16


def pack_beamprops(objs):
    out = np.zeros((len(objs), BEAMPROP_STRIDE))
    for i, o in enumerate(objs):
        for j, f in enumerate(BEAMPROP_FIELDS):
            out[i, j] = getattr(o, f)
    return out


def _unit(v):
    return v / np.sqrt((v * v).sum(-1))[:, None]


def beam_frames(x, conn, vxy):
    """BeamC/BeamLR.update_rotation_matrix (beamc.pyx:158-231)."""
    X = x.reshape(-1, 3)[conn]
    xh = _unit(X[:, 1] - X[:, 0])
    zh = _unit(np.cross(xh, np.broadcast_to(vxy, xh.shape)))
    yh = _unit(np.cross(zh, xh))
    return np.stack([xh, yh, zh], axis=2)


def truss_frames(x, conn):
    """Truss.update_rotation_matrix (truss.pyx:203-272): vxy = cyclic shift of x-hat."""
    X = x.reshape(-1, 3)[conn]
    xh = _unit(X[:, 1] - X[:, 0])
    vxy = np.stack([xh[:, 1], xh[:, 2], xh[:, 0]], 1)
    zh = _unit(np.cross(xh, vxy))
    yh = _unit(np.cross(zh, xh))
    return np.stack([xh, yh, zh], axis=2)


def spring_frames(axes):
    """Spring.update_rotation_matrix(xi,xj,xk,vxyi,vxyj,vxyk) (spring.pyx:147-206)."""
    axes = np.asarray(axes, float)
    xh = _unit(axes[:, :3])
    zh = _unit(np.cross(xh, axes[:, 3:]))
    yh = _unit(np.cross(zh, xh))
    return np.stack([xh, yh, zh], axis=2)


def _sym(ne, terms):
    K = np.zeros((ne, 12, 12))
    for (i, j), v in terms.items():
        K[:, i, j] = v
        K[:, j, i] = v
    return K


def _p(pe):
    return {f: pe[:, i] for i, f in enumerate(BEAMPROP_FIELDS)}


def _ab(p, L):
    ay = 12 * p["E"] * p["Izz"] / (p["G"] * p["A"] * L ** 2)
    az = 12 * p["E"] * p["Iyy"] / (p["G"] * p["A"] * L ** 2)
    return ay, az, 1 / (1. - ay), 1 / (1. - az)


def beamc_Ke(L, pe):
    p = _p(pe)
    A, E, G, Iyy, Izz, Iyz, J, Ay, Az = [p[k] for k in BEAMPROP_FIELDS[:9]]
    ay, az, by, bz = _ab(p, L)
    Ky = by ** 2 * (A * G * L ** 2 * ay ** 2 + 12 * E * Izz)
    Kz = bz ** 2 * (A * G * L ** 2 * az ** 2 + 12 * E * Iyy)
    Kyz = E * Iyz * by * bz
    Sy = Az * G * ay * by
    Sz = Ay * G * az * bz
    cz = Az * E * bz * (1 - az) / L
    cy = Ay * E * by * (ay - 1) / L
    t = {}
    t[0, 0] = t[6, 6] = A * E / L
    t[0, 6] = -A * E / L
    t[0, 4] = t[6, 10] = cz
    t[0, 10] = t[4, 6] = -cz
    t[0, 5] = t[6, 11] = cy
    t[0, 11] = t[5, 6] = -cy
    t[1, 1] = t[7, 7] = Ky / L ** 3
    t[1, 7] = -Ky / L ** 3
    t[1, 5] = t[1, 11] = Ky / (2 * L ** 2)
    t[5, 7] = t[7, 11] = -Ky / (2 * L ** 2)
    t[2, 2] = t[8, 8] = Kz / L ** 3
    t[2, 8] = -Kz / L ** 3
    t[2, 4] = t[2, 10] = -Kz / (2 * L ** 2)
    t[4, 8] = t[8, 10] = Kz / (2 * L ** 2)
    t[1, 2] = t[7, 8] = 12 * Kyz / L ** 3
    t[1, 8] = t[2, 7] = -12 * Kyz / L ** 3
    t[1, 4] = t[1, 10] = t[5, 8] = t[8, 11] = -6 * Kyz / L ** 2
    t[2, 5] = t[2, 11] = t[4, 7] = t[7, 10] = 6 * Kyz / L ** 2
    t[3, 3] = t[9, 9] = G * J / L
    t[3, 9] = -G * J / L
    t[1, 3] = t[7, 9] = Sy / L
    t[1, 9] = t[3, 7] = -Sy / L
    t[3, 5] = t[3, 11] = Sy / 2
    t[5, 9] = t[9, 11] = -Sy / 2
    t[2, 3] = t[8, 9] = -Sz / L
    t[2, 9] = t[3, 8] = Sz / L
    t[3, 4] = t[3, 10] = Sz / 2
    t[4, 9] = t[9, 10] = -Sz / 2
    t[4, 4] = t[10, 10] = bz ** 2 * (A * G * L ** 2 * az ** 2 / 4 + E * Iyy * az ** 2 - 2 * E * Iyy * az + 4 * E * Iyy) / L
    t[4, 10] = bz ** 2 * (A * G * L ** 2 * az ** 2 / 4 - E * Iyy * az ** 2 + 2 * E * Iyy * az + 2 * E * Iyy) / L
    t[5, 5] = t[11, 11] = by ** 2 * (A * G * L ** 2 * ay ** 2 / 4 + E * Izz * ay ** 2 - 2 * E * Izz * ay + 4 * E * Izz) / L
    t[5, 11] = by ** 2 * (A * G * L ** 2 * ay ** 2 / 4 - E * Izz * ay ** 2 + 2 * E * Izz * ay + 2 * E * Izz) / L
    t[4, 5] = t[10, 11] = Kyz * (-ay * az + ay + az - 4) / L
    t[4, 11] = t[5, 10] = Kyz * (ay * az - ay - az - 2) / L
    return _sym(L.size, t)


def beamc_KGe(L, pe, ue):
    p = _p(pe)
    ay, az, by, bz = _ab(p, L)
    N = p["A"] * p["E"] * (ue[:, 6] - ue[:, 0]) / L
    t = {}
    v = N * by ** 2 * (5 * ay ** 2 - 10 * ay + 6) / (5 * L)
    t[1, 1] = t[7, 7] = v
    t[1, 7] = -v
    v = N * bz ** 2 * (5 * az ** 2 - 10 * az + 6) / (5 * L)
    t[2, 2] = t[8, 8] = v
    t[2, 8] = -v
    v = N * by * bz * (5 * ay * az - 5 * ay - 5 * az + 6) / (5 * L)
    t[1, 2] = t[7, 8] = v
    t[1, 8] = t[2, 7] = -v
    v = N * by ** 2 / 10
    t[1, 5] = t[1, 11] = v
    t[5, 7] = t[7, 11] = -v
    v = -N * bz ** 2 / 10
    t[2, 4] = t[2, 10] = v
    t[4, 8] = t[8, 10] = -v
    v = -N * by * bz / 10
    t[1, 4] = t[1, 10] = v
    t[4, 7] = t[7, 10] = -v
    v = N * by * bz / 10
    t[2, 5] = t[2, 11] = v
    t[5, 8] = t[8, 11] = -v
    t[4, 4] = t[10, 10] = L * N * bz ** 2 * (5 * az ** 2 - 10 * az + 8) / 60
    t[5, 5] = t[11, 11] = L * N * by ** 2 * (5 * ay ** 2 - 10 * ay + 8) / 60
    t[4, 10] = L * N * bz ** 2 * (-5 * az ** 2 + 10 * az - 2) / 60
    t[5, 11] = L * N * by ** 2 * (-5 * ay ** 2 + 10 * ay - 2) / 60
    t[4, 5] = t[10, 11] = L * N * by * bz * (-5 * ay * az + 5 * ay + 5 * az - 8) / 60
    t[4, 11] = t[5, 10] = L * N * by * bz * (5 * ay * az - 5 * ay - 5 * az + 2) / 60
    return _sym(L.size, t)


def beamc_Me(L, pe, mtype):
    p = _p(pe)
    ay, az, by, bz = _ab(p, L)
    r0, ry, rz, ry2, rz2, ryz = [p[k] for k in BEAMPROP_FIELDS[9:]]
    t = {}
    if mtype == 1:
        d = [L * r0 / 2, L * by ** 2 * r0 * (ay - 1) ** 2 / 2, L * bz ** 2 * r0 * (az - 1) ** 2 / 2,
             L * (ry2 + rz2) / 2, L * bz ** 2 * rz2 * (az - 1) ** 2 / 2, L * by ** 2 * ry2 * (ay - 1) ** 2 / 2]
        for i in range(6):
            t[i, i] = t[i + 6, i + 6] = d[i]
        return _sym(L.size, t)
    L2 = L ** 2
    t[0, 0] = t[6, 6] = L * r0 / 3
    t[0, 6] = L * r0 / 6
    t[0, 1] = t[1, 6] = by * ry / 2
    t[0, 7] = t[6, 7] = -by * ry / 2
    t[0, 2] = t[2, 6] = bz * rz / 2
    t[0, 8] = t[6, 8] = -bz * rz / 2
    t[0, 4] = t[6, 10] = L * bz * rz * (1 - 4 * az) / 12
    t[0, 5] = t[6, 11] = L * by * ry * (4 * ay - 1) / 12
    t[0, 10] = t[4, 6] = -L * bz * rz * (2 * az + 1) / 12
    t[0, 11] = t[5, 6] = L * by * ry * (2 * ay + 1) / 12
    t[1, 1] = t[7, 7] = by ** 2 * (70 * L2 * ay ** 2 * r0 - 147 * L2 * ay * r0 + 78 * L2 * r0 + 252 * ry2) / (210 * L)
    t[1, 7] = by ** 2 * (35 * L2 * ay ** 2 * r0 - 63 * L2 * ay * r0 + 27 * L2 * r0 - 252 * ry2) / (210 * L)
    t[2, 2] = t[8, 8] = bz ** 2 * (70 * L2 * az ** 2 * r0 - 147 * L2 * az * r0 + 78 * L2 * r0 + 252 * rz2) / (210 * L)
    t[2, 8] = bz ** 2 * (35 * L2 * az ** 2 * r0 - 63 * L2 * az * r0 + 27 * L2 * r0 - 252 * rz2) / (210 * L)
    t[1, 2] = t[7, 8] = 6 * by * bz * ryz / (5 * L)
    t[1, 8] = t[2, 7] = -6 * by * bz * ryz / (5 * L)
    t[1, 3] = t[7, 9] = L * by * rz * (20 * ay - 21) / 60
    t[1, 9] = t[3, 7] = L * by * rz * (10 * ay - 9) / 60
    t[2, 3] = t[8, 9] = L * bz * ry * (21 - 20 * az) / 60
    t[2, 9] = t[3, 8] = L * bz * ry * (9 - 10 * az) / 60
    t[1, 4] = t[1, 10] = -by * bz * ryz * (5 * az + 1) / 10
    t[4, 7] = t[7, 10] = by * bz * ryz * (5 * az + 1) / 10
    t[2, 5] = t[2, 11] = by * bz * ryz * (5 * ay + 1) / 10
    t[5, 8] = t[8, 11] = -by * bz * ryz * (5 * ay + 1) / 10
    t[1, 5] = by ** 2 * (35 * L2 * ay ** 2 * r0 - 77 * L2 * ay * r0 + 44 * L2 * r0 + 420 * ay * ry2 + 84 * ry2) / 840
    t[7, 11] = -t[1, 5]
    t[1, 11] = by ** 2 * (-35 * L2 * ay ** 2 * r0 + 63 * L2 * ay * r0 - 26 * L2 * r0 + 420 * ay * ry2 + 84 * ry2) / 840
    t[5, 7] = -t[1, 11]
    t[2, 4] = bz ** 2 * (-35 * L2 * az ** 2 * r0 + 77 * L2 * az * r0 - 44 * L2 * r0 - 420 * az * rz2 - 84 * rz2) / 840
    t[8, 10] = -t[2, 4]
    t[2, 10] = bz ** 2 * (35 * L2 * az ** 2 * r0 - 63 * L2 * az * r0 + 26 * L2 * r0 - 420 * az * rz2 - 84 * rz2) / 840
    t[4, 8] = -t[2, 10]
    t[3, 3] = t[9, 9] = L * (ry2 + rz2) / 3
    t[3, 9] = L * (ry2 + rz2) / 6
    t[3, 4] = L2 * bz * ry * (5 * az - 6) / 120
    t[9, 10] = -t[3, 4]
    t[3, 5] = L2 * by * rz * (5 * ay - 6) / 120
    t[9, 11] = -t[3, 5]
    t[3, 10] = L2 * bz * ry * (4 - 5 * az) / 120
    t[4, 9] = -t[3, 10]
    t[3, 11] = L2 * by * rz * (4 - 5 * ay) / 120
    t[5, 9] = -t[3, 11]
    t[4, 4] = t[10, 10] = L * bz ** 2 * (7 * L2 * az ** 2 * r0 - 14 * L2 * az * r0 + 8 * L2 * r0 + 280 * az ** 2 * rz2 - 140 * az * rz2 + 112 * rz2) / 840
    t[5, 5] = t[11, 11] = L * by ** 2 * (7 * L2 * ay ** 2 * r0 - 14 * L2 * ay * r0 + 8 * L2 * r0 + 280 * ay ** 2 * ry2 - 140 * ay * ry2 + 112 * ry2) / 840
    t[4, 10] = L * bz ** 2 * (-7 * L2 * az ** 2 * r0 + 14 * L2 * az * r0 - 6 * L2 * r0 + 140 * az ** 2 * rz2 + 140 * az * rz2 - 28 * rz2) / 840
    t[5, 11] = L * by ** 2 * (-7 * L2 * ay ** 2 * r0 + 14 * L2 * ay * r0 - 6 * L2 * r0 + 140 * ay ** 2 * ry2 + 140 * ay * ry2 - 28 * ry2) / 840
    t[4, 5] = t[10, 11] = L * by * bz * ryz * (-20 * ay * az + 5 * ay + 5 * az - 8) / 60
    t[4, 11] = t[5, 10] = L * by * bz * ryz * (-10 * ay * az - 5 * ay - 5 * az + 2) / 60
    return _sym(L.size, t)


def beamlr_Ke(L, pe):
    p = _p(pe)
    A, E, G, Iyy, Izz, Iyz, J, Ay, Az = [p[k] for k in BEAMPROP_FIELDS[:9]]
    t = {}
    t[0, 0] = t[6, 6] = E * A / L
    t[0, 6] = -E * A / L
    t[0, 4] = t[6, 10] = E * Az / L
    t[0, 10] = t[4, 6] = -E * Az / L
    t[0, 5] = t[6, 11] = -E * Ay / L
    t[0, 11] = t[5, 6] = E * Ay / L
    t[1, 1] = t[7, 7] = t[2, 2] = t[8, 8] = G * A / L
    t[1, 7] = t[2, 8] = -G * A / L
    t[1, 3] = t[7, 9] = -G * Az / L
    t[1, 9] = t[3, 7] = G * Az / L
    t[1, 5] = t[1, 11] = G * A / 2
    t[5, 7] = t[7, 11] = -G * A / 2
    t[2, 3] = t[8, 9] = G * Ay / L
    t[2, 9] = t[3, 8] = -G * Ay / L
    t[2, 4] = t[2, 10] = -G * A / 2
    t[4, 8] = t[8, 10] = G * A / 2
    t[3, 3] = t[9, 9] = G * J / L
    t[3, 9] = -G * J / L
    t[3, 4] = t[3, 10] = -G * Ay / 2
    t[4, 9] = t[9, 10] = G * Ay / 2
    t[3, 5] = t[3, 11] = -G * Az / 2
    t[5, 9] = t[9, 11] = G * Az / 2
    t[4, 4] = t[10, 10] = G * A * L / 4 + E * Iyy / L
    t[4, 10] = G * A * L / 4 - E * Iyy / L
    t[5, 5] = t[11, 11] = G * A * L / 4 + E * Izz / L
    t[5, 11] = G * A * L / 4 - E * Izz / L
    t[4, 5] = t[10, 11] = -E * Iyz / L
    t[4, 11] = t[5, 10] = E * Iyz / L
    return _sym(L.size, t)


def beamlr_KGe(L, pe, ue):
    p = _p(pe)
    N = p["A"] * p["E"] * (ue[:, 6] - ue[:, 0]) / L
    third, sixth = 0.333333333333333, 0.166666666666667   # literal constants, beamlr.pyx:1391
    t = {}
    t[4, 4] = t[5, 5] = t[10, 10] = t[11, 11] = third * L * N
    t[4, 5] = t[10, 11] = -third * L * N
    t[4, 10] = t[5, 11] = sixth * L * N
    t[4, 11] = t[5, 10] = -sixth * L * N
    return _sym(L.size, t)


def _mb(p):
    r0, ry, rz, ry2, rz2, ryz = [p[k] for k in BEAMPROP_FIELDS[9:]]
    ne = r0.size
    mb = np.zeros((ne, 6, 6))
    for d in range(3):
        mb[:, d, d] = r0
    mb[:, 0, 4] = mb[:, 4, 0] = rz
    mb[:, 0, 5] = mb[:, 5, 0] = -ry
    mb[:, 1, 3] = mb[:, 3, 1] = -rz
    mb[:, 2, 3] = mb[:, 3, 2] = ry
    mb[:, 3, 3] = ry2 + rz2
    mb[:, 4, 4] = rz2
    mb[:, 5, 5] = ry2
    mb[:, 4, 5] = mb[:, 5, 4] = -ryz
    return mb


def beamlr_Me(L, pe, mtype, truss=False):
    p = _p(pe)
    mb = _mb(p)
    if truss:   # truss.pyx:894-1800: ry/rz rows and columns carry no inertia
        mb[:, 4:, :] = 0
        mb[:, :, 4:] = 0
    ne = L.size
    M = np.zeros((ne, 12, 12))
    if mtype == 0:
        for a in range(2):
            for b in range(2):
                M[:, 6 * a:6 * a + 6, 6 * b:6 * b + 6] = (L / 3 if a == b else L / 6)[:, None, None] * mb
    else:
        d = np.einsum("eii->ei", mb)
        for a in range(2):
            for i in range(6):
                M[:, 6 * a + i, 6 * a + i] = L / 2 * d[:, i]
    return M


def truss_Ke(L, pe):
    p = _p(pe)
    t = {}
    t[0, 0] = t[6, 6] = p["E"] * p["A"] / L
    t[0, 6] = -p["E"] * p["A"] / L
    t[3, 3] = t[9, 9] = p["G"] * p["J"] / L
    t[3, 9] = -p["G"] * p["J"] / L
    return _sym(L.size, t)


def spring_Ke(k):
    ne = k.shape[0]
    K = np.zeros((ne, 12, 12))
    for i in range(6):
        K[:, i, i] = K[:, i + 6, i + 6] = k[:, i]
        K[:, i, i + 6] = K[:, i + 6, i] = -k[:, i]
    return K


SIZES = {"beamc": dict(KC0=144, KG=144, M=144), "beamlr": dict(KC0=144, KG=36, M=144),
         "truss": dict(KC0=72, KG=0, M=144), "spring": dict(KC0=72, KG=0, M=0)}


def run(case, what, state=False):
    from .shells import local_ue, local_xe
    kind = case["kind"]
    conn = np.asarray(case["conn"], np.int64)
    ne = conn.shape[0]
    sz = SIZES[kind]
    out = {}
    if kind == "spring":
        R = spring_frames(case["axes"])
        pe = None
        L = None
    else:
        x = np.asarray(case["x"], float)
        R = truss_frames(x, conn) if kind == "truss" else beam_frames(x, conn, np.asarray(case["vxy"], float))
        xe = local_xe(R, x, conn)
        L = np.sqrt(((xe[:, 1] - xe[:, 0]) ** 2).sum(1))
        pid = case.get("prop_id")
        pe = np.asarray(case["props"], float)[np.zeros(ne, int) if pid is None else np.asarray(pid)]
    ue = local_ue(R, np.asarray(case["u"], float), conn) if case.get("u") is not None else None
    if kind == "beamc":
        Ke = beamc_Ke(L, pe)
    elif kind == "beamlr":
        Ke = beamlr_Ke(L, pe)
    elif kind == "truss":
        Ke = truss_Ke(L, pe)
    else:
        Ke = spring_Ke(np.asarray(case["k"], float))
    kc0_mask = "full" if kind in ("beamc", "beamlr") else "d18"
    if "KC0" in what:
        out["KC0"] = list(coo.coo_blocks(coo.to_global(Ke, R), conn, kc0_mask, sz["KC0"]))
    if "KG" in what and sz["KG"]:
        if kind == "beamc":
            out["KG"] = list(coo.coo_blocks(coo.to_global(beamc_KGe(L, pe, ue), R), conn, "full", 144))
        else:
            out["KG"] = list(coo.coo_blocks(coo.to_global(beamlr_KGe(L, pe, ue), R), conn, "rr", 36))
    for mt in (0, 1):
        if "M%d" % mt in what and sz["M"]:
            Me = beamc_Me(L, pe, mt) if kind == "beamc" else beamlr_Me(L, pe, mt, truss=(kind == "truss"))
            out["M%d" % mt] = list(coo.coo_blocks(coo.to_global(Me, R), conn, "full" if mt == 0 else "d18",
                                                  sz["M"], diag_pairs_only=(mt == 1)))
    if "fint" in what:
        fe = np.einsum("eij,ej->ei", Ke, ue)
        out["fint"] = coo.scatter_fint(fe, R, conn, case["ndof"])
    if state:
        out.update(R=R, geo=L if L is not None else np.zeros(ne))
        if kind != "spring":
            out["xe"] = xe
    return out
