"""GPU: the fused node-centric evaluate+assemble kernel against the reference goldens (COO values)
and scipy's sum-duplicates of the reference COO (CSR values)."""
import numpy as np
import pytest

from oracle import driver
from tests import cases, util

pytestmark = pytest.mark.gpu


def _check(case, ref, mtype=0, stress=None, node_range=None):
    import scipy.sparse as sp
    from pyfe3d_b200.batch import AssemblyPlan
    b = util.batch_from_case(case)
    n = case["ndof"]
    nn = n // 6
    plan = AssemblyPlan("KC0", nn, [b], node_range=node_range)
    kw = dict(KC0=True, M=True, mtype=mtype, indices=True)
    if stress is None:
        kw["KG"] = True
    else:
        kw["KG_given_stress"] = stress
    coo, csr = plan.evaluate_assemble(**kw)
    ne = case["conn"].shape[0]
    keys = {"KC0": "KC0", "KG": "KG" if stress is None else "KGs", "M": "M%d" % mtype}
    lo, hi = (0, nn) if node_range is None else node_range
    for name, rk in keys.items():
        r, c, v = ref[rk]
        if node_range is None:   # with a shard only the owned nodes' slabs are written
            assert np.array_equal(coo[name].r.cpu().numpy(), r) and np.array_equal(coo[name].c.cpu().numpy(), c)
            assert util.block_relerr(coo[name].v.cpu().numpy(), v, ne) <= util.TOL_VALUES, name
        p = AssemblyPlan(name, nn, [b], node_range=node_range, mtype=mtype)
        A = p.to_scipy(csr[name])
        S = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
        S.sum_duplicates()
        S.sort_indices()
        S = S[6 * lo:6 * hi]
        assert np.array_equal(A.indptr, S.indptr) and np.array_equal(A.indices, S.indices), name
        assert np.abs(A.data - S.data).max() <= util.TOL_CSR * np.abs(S.data).max(), name


@pytest.mark.parametrize("name", ["quad4_mesh", "quad4r_mesh", "quad4_soup", "quad4r_soup", "quad4_soup_thick",
                                  "tria3r_mesh", "tria3r_soup", "tria3r_soup_thick", "quad4_xmat_degenerate",
                                  "quad4r_xmat_degenerate", "tria3r_xmat_degenerate"])
@pytest.mark.parametrize("mtype", [0, 1, 2])
def test_fused_matches_reference_golden(name, mtype):
    case, ref = util.load_golden(name)
    _check(case, ref, mtype=mtype)


@pytest.mark.parametrize("name", ["quad4r_mesh", "tria3r_mesh"])
def test_fused_given_stress_and_shard(name):
    case, ref = util.load_golden(name)
    _check(case, ref, stress=case["stress"])
    nn = case["ndof"] // 6
    _check(case, ref, node_range=(nn // 3, nn - 2))


@pytest.mark.parametrize("kind", ["quad4", "quad4r", "tria3r"])
def test_fused_large_mesh_matches_oracle_and_twopass(kind):
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    case = cases.shell_mesh(kind, 61, 47, seed=404)
    want = driver.run(case, what=("KC0", "KG", "M0"))
    _check(case, want)
    b = util.batch_from_case(case)
    nn = case["ndof"] // 6
    plan = AssemblyPlan("KC0", nn, [b])
    coo, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True)
    two = b.evaluate(KC0=True, KG=True, M=True)
    for m in ("KC0", "KG", "M"):
        assert util.block_relerr(coo[m].v.cpu().numpy(), two[m].v.cpu().numpy(), case["conn"].shape[0]) <= 1e-13
        ref_csr = AssemblyPlan(m, nn, [b]).assemble(two[m].v)
        assert np.abs((csr[m] - ref_csr).cpu().numpy()).max() <= 1e-12 * float(ref_csr.abs().max())
    # CSR only (no COO arrays written)
    _, csr2 = plan.evaluate_assemble(KC0=True, KG=True, M=True, write_coo=False)
    for m in ("KC0", "KG", "M"):
        assert torch.equal(csr2[m], csr[m])


def _fan_case(kind, nquads=6, seed=7):
    """nquads quads around one centre node: the centre has valence nquads (> 4 -> two rounds of node records)."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(0, 2 * np.pi, 2 * nquads, endpoint=False)
    rim = np.stack([np.cos(ang), np.sin(ang), 0.05 * np.sin(3 * ang)], 1) * (1 + 0.1 * rng.uniform(-1, 1, (2 * nquads, 1)))
    X = np.vstack([[0., 0., 0.02], rim]) @ cases.random_rotation(rng).T
    conn = np.array([[0, 1 + 2 * i, 1 + (2 * i + 1) % (2 * nquads), 1 + (2 * i + 2) % (2 * nquads)]
                     for i in range(nquads)], np.int64)
    nn = X.shape[0]
    c = dict(kind=kind, x=X.ravel(), conn=conn, props=cases.random_shellprops(rng, 1), ndof=6 * nn,
             u=1e-4 * rng.normal(size=6 * nn), stress=(1e3, -2e2, 3e2))
    if kind == "quad4r":
        c["hg"] = rng.uniform(0.01, 1., (nquads, 5))
    return c


@pytest.mark.parametrize("kind", ["quad4", "quad4r"])
def test_fused_high_valence_node(kind):
    case = _fan_case(kind)
    want = driver.run(case, what=("KC0", "KG", "M0"))
    _check(case, want)


def _tria_fan_case(ntri=13, seed=9):
    """ntri triangles around one centre node: valence 13 > 10 -> two rounds of triangle node records."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(0, 2 * np.pi, ntri, endpoint=False)
    rim = np.stack([np.cos(ang), np.sin(ang), 0.05 * np.sin(3 * ang)], 1) * (1 + 0.1 * rng.uniform(-1, 1, (ntri, 1)))
    X = np.vstack([[0., 0., 0.02], rim]) @ cases.random_rotation(rng).T
    conn = np.array([[0, 1 + i, 1 + (i + 1) % ntri] for i in range(ntri)], np.int64)
    nn = X.shape[0]
    return dict(kind="tria3r", x=X.ravel(), conn=conn, props=cases.random_shellprops(rng, 1), ndof=6 * nn,
                u=1e-4 * rng.normal(size=6 * nn), stress=(1e3, -2e2, 3e2))


def test_fused_tria_high_valence_node():
    case = _tria_fan_case()
    want = driver.run(case, what=("KC0", "KG", "KGs", "M0"))
    _check(case, want)
    _check(case, want, stress=case["stress"])


def test_evaluate_assemble_falls_back_for_other_kinds():
    """BeamC has no fused kernel: the same call runs evaluation + slab assembly and returns the same outputs."""
    import scipy.sparse as sp
    from pyfe3d_b200.batch import AssemblyPlan
    case, ref = util.load_golden("beamc_chain")
    b = util.batch_from_case(case)
    n = case["ndof"]
    plan = AssemblyPlan("KC0", n // 6, [b])
    coo, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True, mtype=1, indices=True)
    for name, rk in (("KC0", "KC0"), ("KG", "KG"), ("M", "M1")):
        r, c, v = ref[rk]
        assert np.array_equal(coo[name].r.cpu().numpy(), r)
        assert util.block_relerr(coo[name].v.cpu().numpy(), v, case["conn"].shape[0]) <= util.TOL_VALUES
        A = plan._sibling(name, 1).to_scipy(csr[name])
        S = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
        S.sum_duplicates()
        assert abs(A - S).max() <= util.TOL_CSR * np.abs(S.data).max()


@pytest.mark.parametrize("kind", ["quad4", "tria3r"])
def test_host_buffer_step_matches_device_step(kind):
    """pf3_eval_assemble_host (x, u in and CSR values out through HOST buffers) == the device-resident call."""
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    case = cases.shell_mesh(kind, 17, 13, seed=21)
    b = util.batch_from_case(case)
    nn = case["ndof"] // 6
    plan = AssemblyPlan("KC0", nn, [b])
    _, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True, mtype=1, write_coo=False)
    sizes = plan.csr_sizes(1)
    out = {m: torch.empty(sizes[m], dtype=torch.float64).pin_memory() for m in ("KC0", "KG", "M")}
    xh = torch.as_tensor(np.ascontiguousarray(case["x"])).pin_memory()
    uh = torch.as_tensor(np.ascontiguousarray(case["u"])).pin_memory()
    plan.evaluate_assemble_host(xh, uh, out, KC0=True, KG=True, M=True, mtype=1)
    for m in ("KC0", "KG", "M"):
        assert torch.equal(out[m], csr[m].cpu()), m
    # a new displacement field through the host path changes KG only
    out2 = {m: np.empty(sizes[m]) for m in ("KC0", "KG")}
    plan.evaluate_assemble_host(case["x"], 3.0 * case["u"], out2, KC0=True, KG=True)
    assert np.array_equal(out2["KC0"], csr["KC0"].cpu().numpy())
    assert np.abs(out2["KG"] - 3.0 * csr["KG"].cpu().numpy()).max() <= 1e-12 * float(csr["KG"].abs().max()) * 3
    with pytest.raises(ValueError):
        plan.evaluate_assemble_host(torch.as_tensor(case["x"]).cuda(), uh, out, KC0=True)


@pytest.mark.parametrize("kind,mtype", [("quad4", 0), ("quad4r", 2)])
def test_host_buffer_step_pipelined_ranges(kind, mtype):
    """pf3_eval_assemble_host on a mesh large enough (>= 131 072 nodes) for the pipelined path: K1 once, K2 over 8 ranges
    of node pairs, every range's CSR rows copied back on a second stream while the next range is evaluated.  The host
    arrays must equal the device-resident fused result bit for bit (same kernels, same order of operations), with an
    odd node count (a half pair at the end) and the COO arrays still written on the device."""
    import torch
    from pyfe3d_b200 import meshes
    from pyfe3d_b200.batch import AssemblyPlan
    case = meshes.plate_quad4(402, 326, kind=kind)          # 403 x 327 = 131 781 nodes (odd)
    nn = case["ndof"] // 6
    assert nn >= 131072 and nn % 2 == 1
    b = util.batch_from_case(case)
    plan = AssemblyPlan("KC0", nn, [b])
    coo, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True, mtype=mtype)
    sizes = plan.csr_sizes(mtype)
    out = {m: torch.full((sizes[m],), float("nan"), dtype=torch.float64).pin_memory() for m in ("KC0", "KG", "M")}
    xh = torch.as_tensor(np.ascontiguousarray(case["x"])).pin_memory()
    uh = torch.as_tensor(np.ascontiguousarray(case["u"])).pin_memory()
    coo2 = {m: b._alloc(m, False, None) for m in ("KC0", "KG", "M")}
    plan.evaluate_assemble_host(xh, uh, out, KC0=True, KG=True, M=True, mtype=mtype, coo=coo2)
    for m in ("KC0", "KG", "M"):
        assert torch.equal(out[m], csr[m].cpu()), m
        assert torch.equal(coo2[m].v, coo[m].v), m
    # a second call reuses the stream / events, only KG changes with u
    out2 = {m: torch.empty(sizes[m], dtype=torch.float64).pin_memory() for m in ("KC0", "KG")}
    plan.evaluate_assemble_host(xh, 2.0 * uh, out2, KC0=True, KG=True)
    assert torch.equal(out2["KC0"], csr["KC0"].cpu())
    assert (out2["KG"] - 2.0 * csr["KG"].cpu()).abs().max() <= 1e-12 * 2 * float(csr["KG"].abs().max())


def _mixed_reference(cs, keys):
    """scipy sum-duplicates of the oracle's triplets of every batch, per matrix."""
    import scipy.sparse as sp
    n = cs[0]["ndof"]
    outs = [driver.run(c, what=keys) for c in cs]
    mats = {}
    for k in keys:
        r = np.concatenate([o[k][0] for o in outs])
        c = np.concatenate([o[k][1] for o in outs])
        v = np.concatenate([o[k][2] for o in outs])
        S = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
        S.sum_duplicates()
        S.sort_indices()
        mats[k] = S
    return outs, mats


@pytest.mark.parametrize("kind", ["quad4", "quad4r"])
def test_fused_mixed_skin_and_stiffeners(kind):
    """BASELINE config 5 in small: Quad skin + BeamC stiffeners in one KC0 / KG / M each.  The quad share goes through
    the fused kernel into the union layouts (KG and M have 36 entries per block because of the beams), the beams are
    added; COO slices and CSR against the oracle + scipy."""
    import torch
    from pyfe3d_b200 import meshes
    from pyfe3d_b200.batch import AssemblyPlan
    skin, beams = meshes.stiffened_panel(14, 11, nstiff=3)
    skin["kind"] = kind
    rng = np.random.default_rng(3)
    # a beam between two nodes that share no quad (a new block), and one on a node pair of the mesh diagonal
    nny = 12
    extra = np.array([[0, 2 * nny + 5], [3 * nny + 3, 4 * nny + 4]], np.int64)
    beams["conn"] = np.vstack([beams["conn"], extra])
    beams["vxy"] = np.vstack([beams["vxy"], beams["vxy"][:2]])
    cs = [skin, beams]
    bs = [util.batch_from_case(c) for c in cs]
    nn = skin["ndof"] // 6
    keys = ("KC0", "KG", "M0")
    outs, mats = _mixed_reference(cs, keys)
    plan = AssemblyPlan("KC0", nn, bs)
    coo, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True)
    assert not getattr(plan, "_fused_unsupported", False)
    for name, rk in (("KC0", "KC0"), ("KG", "KG"), ("M", "M0")):
        p = plan._sibling(name, 0)
        for g, (c, o) in enumerate(zip(cs, outs)):
            ne = c["conn"].shape[0]
            got = coo[name].v[p.coo_offsets[g]:p.coo_offsets[g] + o[rk][2].size].cpu().numpy()
            assert util.block_relerr(got, o[rk][2], ne) <= util.TOL_VALUES, (name, g)
        A = p.to_scipy(csr[name])
        S = mats[rk]
        # the union-mask layout of a multi-group plan stores explicit zeros where only some kinds have entries
        # (e.g. the quads' KG inside the beams' 6x6 blocks): same matrix, pattern a superset of scipy's
        assert abs(A - S).max() <= util.TOL_CSR * np.abs(S.data).max(), name
        if name == "KC0":
            assert np.array_equal(A.indptr, S.indptr) and np.array_equal(A.indices, S.indices)
    # same outputs as the two-pass path of the same plan
    plan._fused_unsupported = True
    coo2, csr2 = plan.evaluate_assemble(KC0=True, KG=True, M=True)
    for name in ("KC0", "KG", "M"):
        assert float((csr[name] - csr2[name]).abs().max()) <= 1e-12 * float(csr2[name].abs().max()), name
    # a row shard of the same mesh
    lo, hi = nn // 3, nn - 7
    ps = AssemblyPlan("KC0", nn, bs, node_range=(lo, hi))
    _, csrs = ps.evaluate_assemble(KC0=True, KG=True, M=True)
    for name, rk in (("KC0", "KC0"), ("KG", "KG"), ("M", "M0")):
        A = ps._sibling(name, 0).to_scipy(csrs[name])
        S = mats[rk][6 * lo:6 * hi]
        assert abs(A - S).max() <= util.TOL_CSR * np.abs(S.data).max()


def test_fused_mixed_beam_only_nodes():
    """Nodes that only the beams touch (a stiffener sticking out of the skin) still get initialised rows."""
    from pyfe3d_b200 import meshes
    from pyfe3d_b200.batch import AssemblyPlan
    skin, beams = meshes.stiffened_panel(6, 5, nstiff=2)
    nn0 = skin["ndof"] // 6
    rng = np.random.default_rng(8)
    xtra = rng.normal(size=(3, 3)) * 0.1 + np.array([1.2, 0.5, 0.1])
    x = np.concatenate([skin["x"], xtra.ravel()])
    nn = nn0 + 3
    u = np.concatenate([skin["u"], 1e-4 * rng.normal(size=18)])
    for c in (skin, beams):
        c["x"], c["u"], c["ndof"] = x, u, 6 * nn
    beams["conn"] = np.vstack([beams["conn"], [[nn0 - 1, nn0], [nn0, nn0 + 1], [nn0 + 1, nn0 + 2]]]).astype(np.int64)
    beams["vxy"] = np.vstack([beams["vxy"], beams["vxy"][:3]])
    cs = [skin, beams]
    bs = [util.batch_from_case(c) for c in cs]
    keys = ("KC0", "KG", "M0")
    outs, mats = _mixed_reference(cs, keys)
    plan = AssemblyPlan("KC0", nn, bs)
    _, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True)
    assert not getattr(plan, "_fused_unsupported", False)
    for name, rk in (("KC0", "KC0"), ("KG", "KG"), ("M", "M0")):
        A = plan._sibling(name, 0).to_scipy(csr[name])
        S = mats[rk]
        assert np.isfinite(A.data).all()
        assert abs(A - S).max() <= util.TOL_CSR * np.abs(S.data).max(), name


def test_fused_mixed_lumped_beam_mass_takes_two_pass_for_M():
    """mtype = 1 on a skin + stiffener plan: the BeamC group contributes only DIAGONAL node pairs to M
    (lumped beam mass, beamc.pyx:2969), so the M plan has fewer node blocks than the KC0 plan wherever a beam joins two
    nodes that share no quad.  The fused kernel walks the KC0 plan's blocks and must not write such an M: KC0 and KG go
    through the fused group path, M through evaluation + its own plan; all three against the oracle + scipy."""
    from pyfe3d_b200 import _cabi, meshes
    from pyfe3d_b200.batch import AssemblyPlan
    skin, beams = meshes.stiffened_panel(9, 7, nstiff=2)
    nny = 8
    extra = np.array([[0, 2 * nny + 5], [3 * nny + 3, 5 * nny + 1]], np.int64)      # node pairs that share no quad
    beams["conn"] = np.vstack([beams["conn"], extra])
    beams["vxy"] = np.vstack([beams["vxy"], beams["vxy"][:2]])
    cs = [skin, beams]
    bs = [util.batch_from_case(c) for c in cs]
    nn = skin["ndof"] // 6
    outs, mats = _mixed_reference(cs, ("KC0", "KG", "M1"))
    plan = AssemblyPlan("KC0", nn, bs)
    pm = plan._sibling("M", 1)
    assert pm._plan.nblocks < plan._plan.nblocks                    # the hazard is real on this mesh
    coo, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True, mtype=1)
    assert not getattr(plan, "_fused_unsupported", False)           # KC0 / KG still took the fused path
    assert csr["M"].numel() == pm.nnz
    for name, rk, mt in (("KC0", "KC0", 0), ("KG", "KG", 0), ("M", "M1", 1)):
        p = plan._sibling(name, mt)
        for g, (c, o) in enumerate(zip(cs, outs)):
            got = coo[name].v[p.coo_offsets[g]:p.coo_offsets[g] + o[rk][2].size].cpu().numpy()
            assert util.block_relerr(got, o[rk][2], c["conn"].shape[0]) <= util.TOL_VALUES, (name, g)
        A = p.to_scipy(csr[name])
        assert abs(A - mats[rk]).max() <= util.TOL_CSR * np.abs(mats[rk].data).max(), name
    # the C ABI refuses the hazardous combination outright
    with pytest.raises(_cabi.Pf3Error):
        plan._plan.eval_assemble_group(bs[0].cabi_batch(1, (0., 0., 0.), None), 0, _cabi.M, None, None,
                                       _cabi.Coo(0, 0, coo["M"].v.data_ptr(), 0, 0), 0, 0, csr["M"].data_ptr())


def test_plan_rejects_out_of_range_connectivity_and_handles_hub_rows():
    """node ids outside [0, nnodes) are rejected at plan creation; a hub node coupled to 1600 nodes (spider of
    springs) exceeds the shared-memory accumulator of the per-node gather and takes the global-memory variant."""
    import scipy.sparse as sp
    import torch
    from pyfe3d_b200.batch import AssemblyPlan, ElementBatch
    rng = np.random.default_rng(5)
    conn = np.array([[0, 1], [1, 7]], np.int64)
    b = ElementBatch("spring", conn, k=np.ones((2, 6)), axes=np.tile([1., 0, 0, 0, 1., 0], (2, 1)), nnodes=5)
    with pytest.raises(ValueError):
        AssemblyPlan("KC0", 5, [b])
    nsp = 1600          # 1601 blocks x 18 entries x 8 B = 230 kB > the 200 kB shared-memory budget of one warp
    conn = np.stack([np.zeros(nsp, np.int64), np.arange(1, nsp + 1)], 1)
    case = dict(kind="spring", conn=conn, k=10 ** rng.uniform(3, 6, (nsp, 6)),
                axes=np.concatenate([rng.normal(size=(nsp, 3)), rng.normal(size=(nsp, 3))], 1), props=None,
                ndof=6 * (nsp + 1), u=np.zeros(6 * (nsp + 1)), x=np.zeros(3 * (nsp + 1)))
    bb = util.batch_from_case(case)
    coo = bb.update_KC0()
    plan = AssemblyPlan("KC0", nsp + 1, [bb])
    A = plan.to_scipy(plan.assemble(coo.v))
    want = driver.run(case, what=("KC0",))["KC0"]
    S = sp.coo_matrix((want[2], (want[0], want[1])), shape=(case["ndof"],) * 2).tocsr()
    assert abs(A - S).max() <= util.TOL_CSR * np.abs(S.data).max()


@pytest.mark.parametrize("kind,numbering", [("quad4", "structured"), ("quad4", "scattered"), ("quad4r", "scattered"),
                                            ("tria3r", "structured"), ("tria3r", "scattered")])
def test_fused_with_l2_prefetch_table(kind, numbering):
    """Meshes large enough (> 32 chunks of 512 node pairs) for the plan to build its L2 prefetch table (common.cuh:
    kPfChunk): structured numbering (one first-use run per chunk; two for the triangles' split numbering) and a random
    renumbering of nodes AND elements (runs scattered over the whole element range: the table drops them).  The fused
    kernel must give what the two-pass path gives either way -- the prefetch only moves data into L2."""
    import torch
    from pyfe3d_b200.batch import AssemblyPlan
    case = cases.shell_mesh(kind, 211, 187, seed=77)
    nn = case["ndof"] // 6
    if numbering == "scattered":
        rng = np.random.default_rng(5)
        pn = rng.permutation(nn)                       # new position of every node
        inv = np.argsort(pn)
        case["x"] = np.asarray(case["x"]).reshape(nn, 3)[inv].ravel()
        case["u"] = np.asarray(case["u"]).reshape(nn, 6)[inv].ravel()
        pe = rng.permutation(case["conn"].shape[0])
        case["conn"] = np.ascontiguousarray(pn[case["conn"]][pe])
        for k in ("prop_id", "xmat", "hg", "K6ROT", "alpha"):
            if case.get(k) is not None and np.ndim(case[k]) >= 1 and np.shape(case[k])[0] == pe.size:
                case[k] = np.ascontiguousarray(np.asarray(case[k])[pe])
    b = util.batch_from_case(case)
    plan = AssemblyPlan("KC0", nn, [b])
    ne = case["conn"].shape[0]
    coo, csr = plan.evaluate_assemble(KC0=True, KG=True, M=True)
    two = b.evaluate(KC0=True, KG=True, M=True, indices=False)
    for m in ("KC0", "KG", "M"):
        assert util.block_relerr(coo[m].v.cpu().numpy(), two[m].v.cpu().numpy(), ne) <= 1e-13, m
        ref_csr = AssemblyPlan(m, nn, [b]).assemble(two[m].v)
        assert float((csr[m] - ref_csr).abs().max()) <= 1e-12 * float(ref_csr.abs().max()), m
    # and a second step into the same arrays (records, prefetches and evict-first stores of a steady-state loop)
    coo2, csr2 = plan.evaluate_assemble(KC0=True, KG=True, M=True, coo=coo, csr={k: v.clone() for k, v in csr.items()})
    for m in ("KC0", "KG", "M"):
        assert torch.equal(csr2[m], csr[m])
