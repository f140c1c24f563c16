"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerances (BASELINE.json): COO indices bit-exact; KC0v/KGv/Mv/fint
# 1e-12 relative; assembled CSR 1e-11 relative.  "Relative" is measured against the
# largest magnitude of the element block (single entries that are exact
# cancellations have no meaningful entry-wise relative error in either implementation).
TOL_VALUES = 1e-12
TOL_CSR = 1e-11


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("quad4_probe", "aero_", "laminates", "lamination", "sticky_")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case, ref = {}, {}
    for k in z.files:
        if k.startswith("in_"):
            v = z[k]
            case[k[3:]] = v.item() if v.ndim == 0 and k != "in_kind" else v
    case["kind"] = str(z["in_kind"])
    if "stress" in case:
        case["stress"] = tuple(float(t) for t in case["stress"])
    case["ndof"] = int(case["ndof"])
    case.setdefault("props", None)
    for k in z.files:
        if k.startswith("ref_"):
            key = k[4:]
            if key.endswith(("_r", "_c", "_v")) and key[:-2] in ("KC0", "KG", "KGs", "M0", "M1", "M2", "KA_beta", "KA_gamma", "CA"):
                ref.setdefault(key[:-2], [None, None, None])["rcv".index(key[-1])] = z[k]
            else:
                ref[key] = z[k]
    return case, ref


def block_relerr(got, want, ne):
    """max over elements of max|got-want| / max|want| within the element's block."""
    if want.size == 0:
        return 0.0
    g = np.asarray(got, float).reshape(ne, -1)
    w = np.asarray(want, float).reshape(ne, -1)
    sc = np.abs(w).max(1, keepdims=True)
    sc[sc == 0] = 1.0
    return float((np.abs(g - w) / sc).max())


def vec_relerr(got, want):
    s = np.abs(want).max()
    return float(np.abs(np.asarray(got) - want).max() / (s if s > 0 else 1.0))


def compare_outputs(got, ref, ne, tol=TOL_VALUES, keys=None):
    """Assert bit-exact indices and block-relative values for every key in ref."""
    checked = []
    for k, want in ref.items():
        if keys is not None and k not in keys:
            continue
        if k in ("R", "m", "xe", "geo"):
            continue
        assert k in got, "missing output %s" % k
        if k == "fint":
            err = vec_relerr(got[k], want)
            assert err <= tol, "fint rel err %.2e" % err
        else:
            r, c, v = want
            if got[k][0] is not None:
                assert np.array_equal(np.asarray(got[k][0]), r), "%s row indices differ" % k
                assert np.array_equal(np.asarray(got[k][1]), c), "%s col indices differ" % k
            err = block_relerr(got[k][2], v, ne)
            assert err <= tol, "%s value rel err %.2e" % (k, err)
        checked.append(k)
    return checked


# ---------------------------------------------------------------------------- GPU side
def batch_from_case(case, device=None):
    """ElementBatch for a case dict (see oracle/driver.py for the fields)."""
    from pyfe3d_b200.batch import ElementBatch
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


def run_gpu(case, what=("KC0", "KG", "KGs", "M0", "M1", "M2", "fint"), fused=True):
    """Same output dict as oracle.driver.run / ref_loop.run, computed by the CUDA path."""
    import torch
    b = batch_from_case(case)
    kind = case["kind"]
    shell = kind in ("quad4", "quad4r", "tria3r")
    out = {}

    def cpu(coo):
        return [None if coo.r is None else coo.r.cpu().numpy(), None if coo.c is None else coo.c.cpu().numpy(),
                coo.v.cpu().numpy()]
    mts = [mt for mt in (0, 1, 2) if "M%d" % mt in what and b.sizes["M"] and (shell or mt < 2)]
    if fused and "KC0" in what and "KG" in what and b.sizes["KG"] and mts:
        fint = torch.zeros(case["ndof"], dtype=torch.float64, device=b.device) if "fint" in what else None
        res = b.evaluate(KC0=True, KG=True, M=True, mtype=mts[0], fint=fint)
        out["KC0"], out["KG"], out["M%d" % mts[0]] = cpu(res["KC0"]), cpu(res["KG"]), cpu(res["M"])
        if fint is not None:
            out["fint"] = fint.cpu().numpy()
        mts = mts[1:]
    else:
        if "KC0" in what:
            out["KC0"] = cpu(b.update_KC0())
        if "KG" in what and b.sizes["KG"]:
            out["KG"] = cpu(b.update_KG())
        if "fint" in what:
            fint = torch.zeros(case["ndof"], dtype=torch.float64, device=b.device)
            out["fint"] = b.update_fint(fint).cpu().numpy()
    if "KGs" in what and shell:
        out["KGs"] = cpu(b.update_KG_given_stress(*case.get("stress", (0., 0., 0.))))
    for mt in mts:
        out["M%d" % mt] = cpu(b.update_M(mtype=mt))
    return out
