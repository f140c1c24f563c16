"""Drive the compiled reference (``oracle/_ref/pyfe3d``) exactly the way its
own tests do: one Python loop over element objects, each filling its slice of
pre-allocated COO arrays (tests/test_quad4_static_point_load.py:53-78,
tests/test_quad4r_linear_buckling_plate.py:163-166,
tests/test_beamc_natural_freq_curved.py:69-89 in /root/reference).

TEST INFRASTRUCTURE ONLY: used to make/verify golden vectors and as the
``--impl reference`` / ``cpu_baseline`` arm of ``bench.py``.
"""
import os
import sys

import numpy as np

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

SHELL_FIELDS = ["A11", "A12", "A16", "A22", "A26", "A66", "B11", "B12", "B16", "B22", "B26", "B66",
                "D11", "D12", "D16", "D22", "D26", "D66", "E44", "E45", "E55", "scf_k13", "scf_k23", "h",
                "intrho", "intrhoz", "intrhoz2"]
BEAM_FIELDS = ["A", "E", "G", "Iyy", "Izz", "Iyz", "J", "Ay", "Az",
               "intrho", "intrhoy", "intrhoz", "intrhoy2", "intrhoz2", "intrhoyz"]
SHELLS = ("quad4", "quad4r", "tria3r")
BEAMS = ("beamc", "beamlr", "truss")


def available():
    return os.path.isdir(os.path.join(_REF_DIR, "pyfe3d"))


def load():
    """Import the compiled reference as the top-level module ``pyfe3d``."""
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    import pyfe3d
    return pyfe3d


def make_props(kind, table):
    ref = load()
    out = []
    if kind in SHELLS:
        from pyfe3d.shellprop import ShellProp
        for row in table:
            p = ShellProp()
            for j, f in enumerate(SHELL_FIELDS):
                setattr(p, f, float(row[j]))
            out.append(p)
    elif kind in BEAMS:
        from pyfe3d.beamprop import BeamProp
        for row in table:
            p = BeamProp()
            for j, f in enumerate(BEAM_FIELDS):
                setattr(p, f, float(row[j]))
            out.append(p)
    return out


_CLS = {"quad4": "Quad4", "quad4r": "Quad4R", "tria3r": "Tria3R", "beamc": "BeamC",
        "beamlr": "BeamLR", "truss": "Truss", "spring": "Spring"}


def run(case, what=("KC0", "KG", "KGs", "M0", "M1", "M2", "fint"), e0=0, e1=None, state=False):
    """Run the reference loop over elements [e0, e1) of ``case``; returns a dict
    of COO triplets / fint exactly as the reference fills them (init_k = local
    element index * SPARSE_SIZE, arrays zero-initialised)."""
    ref = load()
    kind = case["kind"]
    name = _CLS[kind]
    data = getattr(ref, name + "Data")()
    probe = getattr(ref, name + "Probe")()
    Elem = getattr(ref, name)
    conn = np.asarray(case["conn"], np.int64)
    if e1 is None:
        e1 = conn.shape[0]
    ne = e1 - e0
    nn = conn.shape[1]
    x = np.ascontiguousarray(case.get("x", np.zeros(3)), float)
    u = case.get("u")
    u = None if u is None else np.ascontiguousarray(u, float)
    props = make_props(kind, case["props"]) if kind != "spring" else None
    pid = case.get("prop_id")
    xmat = case.get("xmat")
    K6 = case.get("K6ROT")
    alpha = case.get("alpha")
    hg = case.get("hg")
    stress = case.get("stress")
    vxy = case.get("vxy")
    INT = ref.INT
    out = {}
    sizes = {"KC0": data.KC0_SPARSE_SIZE, "KG": getattr(data, "KG_SPARSE_SIZE", 0),
             "KGs": getattr(data, "KG_SPARSE_SIZE", 0)}
    for mt in (0, 1, 2):
        sizes["M%d" % mt] = getattr(data, "M_SPARSE_SIZE", 0)
    for w in ("KA_beta", "KA_gamma", "CA"):
        sizes[w] = getattr(data, w.upper() + "_SPARSE_SIZE", 0)
    for w in what:
        if w == "fint":
            out["fint"] = np.zeros(case["ndof"])
        elif sizes[w] > 0:
            n = sizes[w] * ne
            out[w] = [np.zeros(n, INT), np.zeros(n, INT), np.zeros(n)]
    if state:
        out["R"] = np.zeros((ne, 3, 3))
        out["m"] = np.zeros((ne, 2, 2))
        out["xe"] = np.zeros((ne, nn, 3))
        out["geo"] = np.zeros(ne)
    for i in range(ne):
        e = e0 + i
        el = Elem(probe)
        for a in range(nn):
            setattr(el, "n%d" % (a + 1), int(conn[e, a]))
            setattr(el, "c%d" % (a + 1), int(6 * conn[e, a]))
        el.init_k_KC0 = i * sizes["KC0"]
        if sizes["KG"]:
            el.init_k_KG = i * sizes["KG"]
        if sizes["M0"]:
            el.init_k_M = i * sizes["M0"]
        prop = props[int(pid[e]) if pid is not None else 0] if props is not None else None
        if kind in SHELLS:
            if K6 is not None:
                el.K6ROT = float(np.broadcast_to(K6, (conn.shape[0],))[e])
            if alpha is not None and kind == "tria3r":
                el.alpha_shear_locking = float(np.broadcast_to(alpha, (conn.shape[0],))[e])
            if xmat is not None:
                el.update_rotation_matrix(x, float(xmat[e, 0]), float(xmat[e, 1]), float(xmat[e, 2]))
            else:
                el.update_rotation_matrix(x)
            el.update_probe_xe(x)
        elif kind in ("beamc", "beamlr"):
            el.update_rotation_matrix(float(vxy[e, 0]), float(vxy[e, 1]), float(vxy[e, 2]), x)
            el.update_probe_xe(x)
        elif kind == "truss":
            el.update_rotation_matrix(x)
            el.update_probe_xe(x)
        else:  # spring
            k6 = case["k"][e]
            el.kxe, el.kye, el.kze, el.krxe, el.krye, el.krze = [float(t) for t in k6]
            ax = case["axes"][e]
            el.update_rotation_matrix(*[float(t) for t in ax])
        if u is not None:
            el.update_probe_ue(u)
        hgk = {}
        if kind == "quad4r" and hg is not None:
            hgk = dict(zip(("hgfactor_u", "hgfactor_v", "hgfactor_w", "hgfactor_rx", "hgfactor_ry"),
                           [float(t) for t in hg[e]]))
        if "KC0" in what:
            if kind == "spring":
                el.update_KC0(*out["KC0"])
            else:
                el.update_KC0(*out["KC0"], prop, **hgk)
        if "KG" in what and "KG" in out:
            el.update_KG(*out["KG"], prop)
        if "KGs" in what and "KGs" in out and kind in SHELLS:
            s = stress if stress is not None else (0., 0., 0.)
            s = [float(np.broadcast_to(t, (conn.shape[0],))[e]) for t in s]
            el.update_KG_given_stress(s[0], s[1], s[2], *out["KGs"])
        for mt in (0, 1, 2):
            w = "M%d" % mt
            if w in what and w in out:
                if mt == 2 and kind not in SHELLS:
                    continue
                el.update_M(*out[w], prop, mtype=mt)
        for w in ("KA_beta", "KA_gamma", "CA"):
            if w in what and w in out:
                setattr(el, "init_k_" + w, i * sizes[w])
                getattr(el, "update_" + w)(*out[w])
        if "fint" in what:
            if kind == "spring":
                el.update_fint(out["fint"])
            else:
                el.update_fint(out["fint"], prop, **hgk)
        if state:
            out["R"][i] = [[el.r11, el.r12, el.r13], [el.r21, el.r22, el.r23], [el.r31, el.r32, el.r33]]
            if kind in SHELLS:
                out["m"][i] = [[el.m11, el.m12], [el.m21, el.m22]]
                out["geo"][i] = el.area
            elif kind != "spring":
                out["geo"][i] = el.length
            if kind != "spring":
                out["xe"][i] = np.asarray(probe.xe).reshape(nn, 3)
    if "KGs" in out and kind not in SHELLS:
        del out["KGs"]
    if "M2" in out and kind not in SHELLS:
        del out["M2"]
    return out
