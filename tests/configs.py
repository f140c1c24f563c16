"""The five BASELINE.json configurations as (mesh case, post-processing) pairs, at sizes a CPU solves in
seconds.  ``build(name)`` returns the case(s); ``scalars(name, mats)`` turns assembled scipy matrices into
the physics numbers the reference's tests assert on (static displacement, natural frequencies, buckling
load).  The same functions are used with matrices from the compiled reference (golden generation,
tests/golden/make_config_scalars.py) and with matrices from the CUDA path (tests/test_gpu_configs.py), so
the comparison at 1e-8 isolates the element/assembly path."""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import eigsh, spsolve

from pyfe3d_b200 import meshes
from pyfe3d_b200.shellprop_utils import isotropic_plate

DOF = 6
NAMES = ("quad4_static", "beamc_freq", "quad4r_cylinder_buckling", "tria3r_freq", "stiffened_panel")


def _plate_nodes(nx, ny, a, b):
    xs, ys = np.meshgrid(np.linspace(0, a, nx), np.linspace(0, b, ny), indexing="ij")
    return np.stack([xs.ravel(), ys.ravel(), np.zeros(nx * ny)], 1)


def build(name):
    if name == "quad4_static":
        # tests/test_quad4_static_point_load.py:15-49: nx=7, ny=11, a=3, b=7, h=0.005, E=200e9, nu=0.3
        nx, ny, a, b = 7, 11, 3., 7.
        X = _plate_nodes(nx, ny, a, b)
        pos = np.arange(nx * ny).reshape(nx, ny)
        conn = np.stack([pos[:-1, :-1].ravel(), pos[1:, :-1].ravel(), pos[1:, 1:].ravel(), pos[:-1, 1:].ravel()], 1)
        prop = isotropic_plate(thickness=0.005, E=200e9, nu=0.3, calc_scf=True)
        return [dict(kind="quad4", x=X.ravel(), conn=conn.astype(np.int64), props=meshes.shellprop_row(prop)[None, :],
                     ndof=DOF * nx * ny, u=np.zeros(DOF * nx * ny), meta=dict(nx=nx, ny=ny, a=a, b=b))]
    if name == "beamc_freq":
        return [meshes.arc_beamc(50)]
    if name == "quad4r_cylinder_buckling":
        return [meshes.cylinder_quad4r(24, 13)]
    if name == "tria3r_freq":
        c = meshes.plate_tria3r(8, 10)
        c["meta"] = dict(nx=9, ny=11, a=0.3, b=0.5)
        return [c]
    if name == "stiffened_panel":
        skin, beams = meshes.stiffened_panel(12, 10, nstiff=3)
        return [skin, beams]
    raise KeyError(name)


def what(name):
    return {"quad4_static": ("KC0",), "beamc_freq": ("KC0", "M0"), "quad4r_cylinder_buckling": ("KC0", "KGs"),
            "tria3r_freq": ("KC0", "M1"), "stiffened_panel": ("KC0", "M0")}[name]


def _free(n, fixed):
    bk = np.zeros(n, bool)
    bk[fixed] = True
    return ~bk


def scalars(name, cases_, mats):
    """mats: dict matrix-key -> scipy CSR (N x N).  Returns a dict of floats."""
    c = cases_[0]
    n = c["ndof"]
    X = c["x"].reshape(-1, 3)
    if name == "quad4_static":
        m = c["meta"]
        x, y = X[:, 0], X[:, 1]
        edge = np.isclose(x, 0.) | np.isclose(x, m["a"]) | np.isclose(y, 0.) | np.isclose(y, m["b"])
        bk = np.zeros(n, bool)
        bk[2::DOF] = edge
        bk[0::DOF] = True
        bk[1::DOF] = True
        bk[5::DOF] = True          # drilling held: direct solve instead of the reference's cg
        bu = ~bk
        f = np.zeros(n)
        f[2::DOF][np.isclose(x, m["a"] / 2) & np.isclose(y, m["b"] / 2)] = 1.
        K = mats["KC0"].tocsc()[bu, :][:, bu]
        u = np.zeros(n)
        u[bu] = spsolve(K, f[bu])
        return {"w_max": float(u[2::DOF].max()), "u_norm": float(np.linalg.norm(u))}
    if name == "beamc_freq":
        bk = np.zeros(n, bool)
        bk[:DOF] = True            # clamped at the first node (tests/test_beamc_natural_freq_curved.py:96-100)
        bu = ~bk
        K = mats["KC0"].tocsc()[bu, :][:, bu]
        M = mats["M0"].tocsc()[bu, :][:, bu]
        vals, _ = eigsh(A=K, M=M, sigma=-1., which="LM", k=3, tol=0)
        om = np.sqrt(np.sort(vals))
        return {"omega1": float(om[0]), "omega2": float(om[1]), "omega3": float(om[2])}
    if name == "quad4r_cylinder_buckling":
        z = X[:, 2]
        ends = np.isclose(z, z.min()) | np.isclose(z, z.max())
        bk = np.zeros(n, bool)
        for d in range(DOF):
            bk[d::DOF] = ends
        bu = ~bk
        K = mats["KC0"].tocsc()[bu, :][:, bu]
        KG = mats["KGs"].tocsc()[bu, :][:, bu]
        # (K + lambda KG) phi = 0, smallest positive lambda
        vals, _ = eigsh(A=KG, M=K, k=4, which="LA", tol=0)   # mu = -1/lambda ... use generalized inverse form
        lam = np.sort(-1. / vals[np.abs(vals) > 0])
        pos = lam[lam > 0]
        neg = -lam[lam < 0]
        return {"lambda_abs_min": float(min(pos.min() if pos.size else np.inf, neg.min() if neg.size else np.inf))}
    if name == "tria3r_freq":
        m = c["meta"]
        x, y = X[:, 0], X[:, 1]
        edge = np.isclose(x, 0.) | np.isclose(x, m["a"]) | np.isclose(y, 0.) | np.isclose(y, m["b"])
        bk = np.zeros(n, bool)
        bk[2::DOF] = edge
        bk[0::DOF] = True
        bk[1::DOF] = True
        bk[5::DOF] = True
        bu = ~bk
        K = mats["KC0"].tocsc()[bu, :][:, bu]
        M = mats["M1"].tocsc()[bu, :][:, bu]
        vals, _ = eigsh(A=K, M=M, sigma=-1., which="LM", k=2, tol=0)
        om = np.sqrt(np.sort(vals))
        return {"omega1": float(om[0]), "omega2": float(om[1])}
    if name == "stiffened_panel":
        Xl = X @ meshes.fixed_rotation(0)       # back to the plate frame (rows of R^T)
        x, y = Xl[:, 0], Xl[:, 1]
        edge = np.isclose(x, x.min()) | np.isclose(x, x.max()) | np.isclose(y, y.min()) | np.isclose(y, y.max())
        bk = np.zeros(n, bool)
        for d in range(DOF):
            bk[d::DOF] = edge
        bu = ~bk
        K = mats["KC0"].tocsc()[bu, :][:, bu]
        M = mats["M0"].tocsc()[bu, :][:, bu]
        normal = meshes.fixed_rotation(0)[:, 2]
        f = np.zeros(n)
        for d in range(3):
            f[d::DOF] = normal[d] * 1e-2
        u = np.zeros(n)
        u[bu] = spsolve(K, f[bu])
        vals, _ = eigsh(A=K, M=M, sigma=-1., which="LM", k=1, tol=0)
        return {"u_norm": float(np.linalg.norm(u)), "omega1": float(np.sqrt(vals[0]))}
    raise KeyError(name)


def assemble_scipy(cases_, outs, key):
    """sum of the groups' COO triplets -> CSR (what the reference scripts do)."""
    n = cases_[0]["ndof"]
    A = sp.csr_matrix((n, n))
    for o in outs:
        if key in o:
            r, c, v = o[key]
            A = A + sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
    return A
