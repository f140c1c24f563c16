"""Oracle entry point: ``run(case, what)`` returns the same dict of COO triplets
and ``fint`` the reference loop (``oracle/ref_loop.py``) produces for ``case``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

A *case* is a plain dict:
  kind      'quad4' | 'quad4r' | 'tria3r' | 'beamc' | 'beamlr' | 'truss' | 'spring'
  x, u      global coordinates [3*nnodes] / DOF vector [6*nnodes] (u optional)
  ndof      6*nnodes
  conn      int64[ne, nn] node positions
  props     float64[nprop, 32] (shells) / [nprop, 16] (beams); prop_id int[ne] optional
  xmat      [ne,3] material direction (shells, optional); K6ROT, alpha, hg[ne,5], stress=(Nxx,Nyy,Nxy)
  vxy       [ne,3] beam orientation vector; spring: k[ne,6], axes[ne,6]
"""
import numpy as np

from . import beams as B
from . import coo
from . import shells as S

SIZES = {
    "quad4": dict(KC0=576, KG=144, M=480), "quad4r": dict(KC0=576, KG=144, M=480),
    "tria3r": dict(KC0=324, KG=81, M=270), "beamc": dict(KC0=144, KG=144, M=144),
    "beamlr": dict(KC0=144, KG=36, M=144), "truss": dict(KC0=72, KG=0, M=144),
    "spring": dict(KC0=72, KG=0, M=0),
}


def _bcast(v, ne, default):
    return np.broadcast_to(np.asarray(default if v is None else v, float), (ne,)).copy()


def run(case, what=("KC0", "KG", "KGs", "M0", "M1", "M2", "fint"), state=False):
    kind = case["kind"]
    if kind in ("quad4", "quad4r", "tria3r"):
        return _run_shell(case, what, state)
    return B.run(case, what, state)


def _run_shell(case, what, state):
    kind = case["kind"]
    conn = np.asarray(case["conn"], np.int64)
    ne, nn = conn.shape
    x = np.asarray(case["x"], float)
    sz = SIZES[kind]
    pid = case.get("prop_id")
    pe = np.asarray(case["props"], float)[np.zeros(ne, int) if pid is None else np.asarray(pid)]
    frames = S.tria_frames if kind == "tria3r" else S.quad_frames
    R, m = frames(x, conn, case.get("xmat"))
    xe = S.local_xe(R, x, conn)
    area = S.tria_area(xe) if kind == "tria3r" else S.quad_area(xe)
    ue = S.local_ue(R, np.asarray(case["u"], float), conn) if case.get("u") is not None else None
    K6 = _bcast(case.get("K6ROT"), ne, 100.)
    alpha = _bcast(case.get("alpha"), ne, 0.7)
    hg = np.ones((ne, 5)) if case.get("hg") is None else np.asarray(case["hg"], float)
    out = {}

    def Ke(for_fint):
        if kind == "quad4":
            return S.quad4_Ke(xe, area, pe, m)
        if kind == "quad4r":
            return S.quad4r_Ke(xe, area, pe, m, K6, hg)
        return S.tria3r_Ke(xe, area, pe, m, K6, alpha, drop_drilling_couplings=not for_fint)

    if "KC0" in what:
        out["KC0"] = list(coo.coo_blocks(coo.to_global(Ke(False), R), conn, "full", sz["KC0"]))
    if "KG" in what:
        Kl = (S.tria_KG_local(xe, area, pe, m, ue=ue) if kind == "tria3r"
              else S.quad_KG_local(xe, pe, m, ue=ue))
        out["KG"] = list(coo.coo_blocks(coo.to_global(Kl, R), conn, "tt", sz["KG"]))
    if "KGs" in what:
        st = case.get("stress") or (0., 0., 0.)
        Kl = (S.tria_KG_local(xe, area, pe, m, stress=st) if kind == "tria3r"
              else S.quad_KG_local(xe, pe, m, stress=st))
        out["KGs"] = list(coo.coo_blocks(coo.to_global(Kl, R), conn, "tt", sz["KG"]))
    if kind != "tria3r" and any(k in what for k in ("KA_beta", "KA_gamma", "CA")):
        locs = dict(zip(("KA_beta", "KA_gamma", "CA"), S.quad_aero_local(xe, R)))
        for k in ("KA_beta", "KA_gamma", "CA"):
            if k in what:
                out[k] = list(coo.coo_blocks(coo.to_global(locs[k], R), conn, "tt", 144))
    for mt in (0, 1, 2):
        if "M%d" % mt in what:
            Ml = S.tria_M_local(xe, area, pe, mt) if kind == "tria3r" else S.quad_M_local(xe, area, pe, mt)
            out["M%d" % mt] = list(coo.coo_blocks(coo.to_global(Ml, R), conn,
                                                  "d18" if mt == 2 else "m30", sz["M"]))
    if "fint" in what:
        fe = np.einsum("eij,ej->ei", Ke(True), ue)
        out["fint"] = coo.scatter_fint(fe, R, conn, case["ndof"])
    if state:
        out.update(R=R, m=m, xe=xe, geo=area)
    return out
