"""The five BASELINE.json configurations: static displacement, natural frequencies and buckling loads from
matrices evaluated+assembled by (a) the numpy oracle [CPU] and (b) the CUDA path [gpu], against the values
obtained from the compiled reference (tests/golden/config_scalars.json), at north_star's 1e-8."""
import json
import os

import numpy as np
import pytest

from oracle import driver
from tests import configs, util

GOLD = json.load(open(os.path.join(util.GOLDEN_DIR, "config_scalars.json")))
RTOL = 1e-8


def _compare(name, got):
    for k, want in GOLD[name].items():
        assert abs(got[k] - want) <= RTOL * abs(want), (name, k, got[k], want)


@pytest.mark.parametrize("name", configs.NAMES)
def test_oracle_reproduces_reference_scalars(name):
    cs = configs.build(name)
    keys = configs.what(name)
    outs = [driver.run(c, what=keys) for c in cs]
    mats = {k: configs.assemble_scipy(cs, outs, k) for k in keys}
    _compare(name, configs.scalars(name, cs, mats))


@pytest.mark.gpu
@pytest.mark.parametrize("name", configs.NAMES)
def test_cuda_reproduces_reference_scalars(name):
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    cs = configs.build(name)
    keys = configs.what(name)
    batches = [util.batch_from_case(c) for c in cs]
    nn = cs[0]["ndof"] // 6
    mats = {}
    for key in keys:
        matrix = {"KC0": "KC0", "KGs": "KG", "M0": "M", "M1": "M"}[key]
        mtype = 1 if key == "M1" else 0
        vals = []
        use = []
        for c, b in zip(cs, batches):
            if b.sizes[matrix] == 0:
                continue
            if key == "KGs":
                coo = b.update_KG_given_stress(*c["stress"], update_KGv_only=1)
            elif matrix == "M":
                coo = b.update_M(mtype=mtype, indices=False)
            else:
                coo = b.update_KC0(update_KC0v_only=1)
            vals.append(coo.v)
            use.append(b)
        plan = AssemblyPlan(matrix, nn, use, mtype=mtype)
        csr = plan.assemble(torch.cat(vals))
        mats[key] = plan.to_scipy(csr)
    _compare(name, configs.scalars(name, cs, mats))


@pytest.mark.gpu
def test_fused_path_static_config():
    """Config 1 through the fused evaluate+assemble call."""
    from pyfe3d_b200.batch import AssemblyPlan
    cs = configs.build("quad4_static")
    b = util.batch_from_case(cs[0])
    plan = AssemblyPlan("KC0", cs[0]["ndof"] // 6, [b])
    _, csr = plan.evaluate_assemble(KC0=True, write_coo=False)
    _compare("quad4_static", configs.scalars("quad4_static", cs, {"KC0": plan.to_scipy(csr["KC0"])}))


@pytest.mark.gpu
def test_static_config_solved_on_device():
    """Config 1 end to end on the GPU: fused evaluate+assemble, boundary conditions as a DOF mask, Jacobi-CG with
    the masked device SpMV; w_max against the reference-derived value at 1e-8."""
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    from pyfe3d_b200.solve import cg_solve, masked_spmv
    cs = configs.build("quad4_static")
    c = cs[0]
    b = util.batch_from_case(c)
    n = c["ndof"]
    plan = AssemblyPlan("KC0", n // 6, [b])
    _, csr = plan.evaluate_assemble(KC0=True, write_coo=False)
    indptr, indices = plan.pattern()
    m = c["meta"]
    X = c["x"].reshape(-1, 3)
    x, y = X[:, 0], X[:, 1]
    edge = np.isclose(x, 0.) | np.isclose(x, m["a"]) | np.isclose(y, 0.) | np.isclose(y, m["b"])
    bk = np.zeros(n, bool)
    bk[2::6] = edge
    bk[0::6] = True
    bk[1::6] = True
    bk[5::6] = True
    f = np.zeros(n)
    f[2::6][np.isclose(x, m["a"] / 2) & np.isclose(y, m["b"] / 2)] = 1.
    free = torch.as_tensor((~bk).astype(np.uint8)).cuda()
    u, info = cg_solve(indptr, indices, csr["KC0"], torch.as_tensor(f).cuda(), free=free, rtol=1e-13)
    assert info > 0
    w_max = float(u[2::6].max())
    want = GOLD["quad4_static"]["w_max"]
    assert abs(w_max - want) <= RTOL * abs(want)
    # masked SpMV == scipy on the extracted sub-matrix
    A = plan.to_scipy(csr["KC0"]).tocsc()
    bu = ~bk
    xv = np.random.default_rng(0).normal(size=n)
    yy = masked_spmv(indptr, indices, csr["KC0"], free, torch.as_tensor(xv).cuda()).cpu().numpy()
    ref = np.zeros(n)
    ref[bu] = A[bu, :][:, bu] @ xv[bu]
    assert np.abs(yy - ref).max() <= 1e-12 * np.abs(ref).max()
