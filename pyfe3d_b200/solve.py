"""Device-side consumers of the assembled CSR matrix (first step): masked SpMV for the
boundary-condition partition and a Jacobi-preconditioned conjugate gradient, so that a static solve never
leaves the GPU.  What it replaces in a reference script (tests/test_quad4_static_point_load.py:84-104)::

    KC0uu = KC0[bu, :][:, bu]                 ->  u = cg_solve(indptr, indices, vals, fext, free=bu)
    uu, info = cg(KC0uu, fext[bu], atol=1e-9)

Single device: ``plan_cg_solve`` runs ``pf3_plan_cg`` -- the whole iteration in native kernels (block SpMV + three fused
vector kernels, deterministic reductions, convergence flag on the device).  Row-sharded over several GPUs it keeps the
loop here, because the halo of the search direction is exchanged between ranks every iteration (grouped NCCL
send / recv between the ranks whose row blocks meet: ``halo_plan`` / ``halo_exchange``).
``compact_csr`` / ``plan_compact`` return ``K[bu, :][:, bu]`` itself as a CSR matrix for callers that hand Kuu to scipy's
``spsolve`` / ``eigsh``."""
import os

import torch

from .batch import _dev, _ptr, context


def masked_spmv(indptr, indices, vals, free, x, out=None):
    """y = P A P x with P = diag(free)."""
    n = indptr.numel() - 1
    if out is None:
        out = torch.empty(n, dtype=torch.float64, device=vals.device)
    context(vals.device).spmv_csr_masked(n, _ptr(indptr), _ptr(indices), _ptr(vals), _ptr(free), _ptr(x), _ptr(out))
    return out


def diagonal(indptr, indices, vals, row0=0):
    n = indptr.numel() - 1
    d = torch.empty(n, dtype=torch.float64, device=vals.device)
    context(vals.device).csr_diagonal(n, _ptr(indptr), _ptr(indices), _ptr(vals), row0, _ptr(d))
    return d


def cg_solve(indptr, indices, vals, b, free=None, rtol=1e-12, maxiter=None, x0=None):
    """Solve (P A P) x = P b for symmetric positive definite A[free, free] with Jacobi-preconditioned CG.
    Returns (x, info): x is zero on constrained DOFs; info = iterations used (negative: not converged)."""
    dev = vals.device
    n = indptr.numel() - 1
    if free is None:
        free = torch.ones(n, dtype=torch.uint8, device=dev)
    free = _dev(free, torch.uint8, dev)
    b = _dev(b, torch.float64, dev) * free
    d = diagonal(indptr, indices, vals)
    minv = torch.where((free > 0) & (d != 0), 1.0 / d, torch.zeros_like(d))
    x = torch.zeros(n, dtype=torch.float64, device=dev) if x0 is None else _dev(x0, torch.float64, dev) * free
    r = b - masked_spmv(indptr, indices, vals, free, x)
    z = minv * r
    p = z.clone()
    rz = torch.dot(r, z)
    bnorm = torch.linalg.vector_norm(b)
    if float(bnorm) == 0.0:
        return x, 0
    maxiter = maxiter or 10 * n
    ap = torch.empty_like(x)
    tol = rtol * float(bnorm)
    if float(torch.linalg.vector_norm(r)) <= tol or float(rz) == 0.0:
        return x, 0
    for it in range(1, maxiter + 1):
        masked_spmv(indptr, indices, vals, free, p, out=ap)
        pap = torch.dot(p, ap)
        if it % 8 == 1 and not float(pap) > 0.0:           # breakdown: not positive definite on the free dofs
            return x, -it
        alpha = rz / pap
        x += alpha * p
        r -= alpha * ap
        z = minv * r
        rz_new = torch.dot(r, z)
        if (it % 8 == 0 or it == maxiter) and (float(torch.linalg.vector_norm(r)) <= tol or float(rz_new) == 0.0):
            return x, it
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, -maxiter


def plan_cg_solve(plan, vals, b, free=None, rtol=1e-12, maxiter=None, x0=None, group=None, extra=(), check_every=16,
                  graph=None):
    """Jacobi-preconditioned CG on the CSR values of a structured ``AssemblyPlan`` through its block SpMV
    (``pf3_plan_spmv``: no per-entry indices, 8.2 B per nonzero).

    The operator is ``A = vals + sum(c * v for plan_i, v, c in extra)`` applied as a sum of block SpMVs -- e.g.
    ``K - sigma M`` with the two matrices kept in their own layouts (``extra=[(plan_M, M, -sigma)]``); every plan
    must own the same rows.

    Single GPU: the plan owns every row.  Multi GPU (``group`` = a torch.distributed process group, one rank per
    GPU): every rank owns the row block of its plan (``node_range``) and holds vectors of its own rows; the search
    direction is the only vector a rank needs beyond its rows, and only at the columns its elements touch, so each
    iteration does ONE point-to-point halo exchange of p (the halo exchange of SURVEY 8(e): grouped NCCL send/recv
    between the ranks whose row blocks meet) and two small all_reduces; the iteration is allocation-free and the host looks at the residual
    once per batch of ``check_every`` iterations.  ``graph=True`` replays each batch as one CUDA graph with the NCCL
    operations captured; it measures the same as eager launches (0.46 against 0.45 ms per iteration on 8 B200 at 24 M
    dofs), so it is off by default.  ``b``/``free``/``x0`` are global [6*nnodes] arrays; returns (x_global, info)."""
    import torch.distributed as dist
    dev = vals.device
    n = 6 * plan.nnodes
    lo, hi = 6 * plan.node_begin, 6 * plan.node_end
    multi = group is not None and dist.get_world_size(group) > 1
    if not multi and (lo != 0 or hi != n):
        raise ValueError("a row-sharded plan needs the process group of the other shards")
    if free is None:
        free = torch.ones(n, dtype=torch.uint8, device=dev)
    free = _dev(free, torch.uint8, dev)
    if not multi:
        x, it, status, _, _ = plan_cg_native(plan, vals, b, free=free, rtol=rtol, maxiter=maxiter, x0=x0, extra=extra)
        return x, (it if status == 0 else -max(it, 1))
    fl = free[lo:hi].to(torch.float64)
    bl = _dev(b, torch.float64, dev)[lo:hi] * fl
    d = plan.diagonal(vals)
    for pe, ve, ce in extra:
        d = d + ce * pe.diagonal(ve)
    minv = torch.where((fl > 0) & (d != 0), 1.0 / d, torch.zeros_like(d))
    tmp = torch.empty(hi - lo, dtype=torch.float64, device=dev) if extra else None

    def matvec(xg, out):
        plan.spmv(vals, xg, free=free, out=out)
        for pe, ve, ce in extra:
            pe.spmv(ve, xg, free=free, out=tmp)
            out.add_(tmp, alpha=ce)
        return out

    # Halo exchange.  A rank's rows read the columns of the nodes its elements touch: [need_lo, need_hi) is that range
    # (from the connectivity), and what lies outside the own rows comes from the ranks that own it -- point-to-point
    # (NCCL send / recv grouped into one launch), straight between the global-length direction vectors: on a banded
    # numbering that is one mesh line of dofs per neighbour instead of an all_gather of the whole vector.
    world, me = dist.get_world_size(group), dist.get_rank(group)
    cmin = min(int(bb.conn.min()) for pp in (plan,) + tuple(e[0] for e in extra) for bb in pp.batches)
    cmax = max(int(bb.conn.max()) for pp in (plan,) + tuple(e[0] for e in extra) for bb in pp.batches)
    mine = torch.tensor([lo, hi, min(6 * cmin, lo), max(6 * cmax + 6, hi)], dtype=torch.int64, device=dev)
    table = torch.empty(4 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(table, mine, group=group)
    table = table.view(world, 4).tolist()
    halo = halo_plan(table, me)
    pg = torch.zeros(n, dtype=torch.float64, device=dev)          # direction, global length: own rows + halo are current

    full_gather = os.environ.get("PF3_CG_EXCHANGE", "halo") == "all_gather"    # A/B switch for scripts/bench_cg_multi.py

    def exchange():
        if full_gather:
            pg.copy_(gather_all(p.clone()))
            return pg
        return halo_exchange(pg, halo, group)

    def allsum(t):
        dist.all_reduce(t, group=group)
        return t

    def gather_all(local):
        sizes = [row[1] - row[0] for row in table]
        out = torch.empty(n, dtype=torch.float64, device=dev)
        if len(set(sizes)) == 1:
            dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        else:
            parts = [torch.empty(sz, dtype=torch.float64, device=dev) for sz in sizes]
            dist.all_gather(parts, local.contiguous(), group=group)
            torch.cat(parts, out=out)
        return out

    p = pg[lo:hi]                                                 # the own slice IS the local direction vector
    x = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
    ap = torch.empty_like(x)
    if x0 is not None:
        x = _dev(x0, torch.float64, dev)[lo:hi] * fl
        p.copy_(x)
        r = bl - matvec(exchange(), ap)
    else:
        r = bl.clone()
    z = minv * r
    p.copy_(z)
    sc = allsum(torch.stack([torch.dot(r, z), torch.dot(r, r), torch.dot(bl, bl)]))
    rz = sc[0].clone()
    bnorm = float(sc[2].sqrt())
    maxiter = maxiter or 10 * n
    if bnorm == 0.0 or float(sc[1].sqrt()) <= rtol * bnorm or float(rz) == 0.0:
        return gather_all(x), 0
    # sc = [p.Ap, new r.z, r.r, current r.z] on the device; the three fused vector kernels of the iteration are native
    # (pf3_cg_shard_dot / _update / _dir, csrc/solve.cu: deterministic reductions), NCCL all-reduces the scalars between
    from . import _cabi
    nl = hi - lo
    sc = torch.zeros(4, dtype=torch.float64, device=dev)
    sc[3] = rz
    work = torch.zeros(_cabi.cg_shard_work_bytes() // 8 + 1, dtype=torch.float64, device=dev)
    del z

    def iterate():
        # one iteration, allocation-free and without host synchronisation
        matvec(exchange(), ap)
        ctx = context(dev)
        ctx.cg_shard_dot(nl, _ptr(p), _ptr(ap), _ptr(sc), _ptr(work))
        allsum(sc[0:1])
        ctx.cg_shard_update(nl, _ptr(p), _ptr(ap), _ptr(minv), _ptr(x), _ptr(r), _ptr(sc), _ptr(work))
        allsum(sc[1:3])                                    # one all_reduce for both scalars
        ctx.cg_shard_dir(nl, _ptr(r), _ptr(minv), _ptr(p), _ptr(sc), _ptr(work))

    def state():
        # (breakdown, converged) from the scalars of the last iteration -- the only host synchronisation of a batch
        t = sc.tolist()
        return (not t[0] > 0.0), (t[2] ** 0.5 <= rtol * bnorm or t[1] == 0.0)

    it, info = 0, -maxiter
    graph_obj = None
    want_graph = bool(graph)
    while it < maxiter:
        k = min(check_every, maxiter - it)
        if want_graph and graph_obj is None and it >= check_every and k == check_every:
            # the first batch ran eagerly (warm-up: NCCL channels, allocator); capture the second one
            try:
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(check_every):
                        iterate()
                graph_obj = g                              # capture does not execute: replay below runs this batch
            except Exception:                              # NCCL / driver without capture support: stay eager
                want_graph = False
                torch.cuda.synchronize(dev)
        if graph_obj is not None and k == check_every:
            graph_obj.replay()
        else:
            for _ in range(k):
                iterate()
        it += k
        broke, conv = state()
        if broke:                                          # not positive definite on the free dofs (or NaN)
            info = -it
            break
        if conv:
            info = it
            break
    return gather_all(x), info


def halo_exchange(vec, halo, group):
    """One halo exchange of the global-length vector ``vec`` (own rows current on entry, own rows + halo on return):
    every (send, recv) pair of ``halo_plan`` as grouped point-to-point operations (one NCCL launch)."""
    import torch.distributed as dist
    ops = []
    for r, (s0, s1), (r0, r1) in halo:
        peer = dist.get_global_rank(group, r) if group is not None and group is not dist.group.WORLD else r
        if s1 > s0:
            ops.append(dist.P2POp(dist.isend, vec[s0:s1], peer, group=group))
        if r1 > r0:
            ops.append(dist.P2POp(dist.irecv, vec[r0:r1], peer, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return vec


def halo_plan(table, me):
    """Who sends what to whom: ``table[r] = (lo, hi, need_lo, need_hi)`` (dof ranges: rows owned by rank r, columns its
    rows read).  Returns [(r, (send_lo, send_hi), (recv_lo, recv_hi))] for every rank r != me with something to move:
    rank `me` sends the part of ITS rows that r reads and receives the part of r's rows that IT reads."""
    lo, hi, need_lo, need_hi = table[me]
    out = []
    for r, (rlo, rhi, rneed_lo, rneed_hi) in enumerate(table):
        if r == me:
            continue
        send = (max(rneed_lo, lo), min(rneed_hi, hi))
        recv = (max(need_lo, rlo), min(need_hi, rhi))
        if send[1] > send[0] or recv[1] > recv[0]:
            out.append((r, send if send[1] > send[0] else (0, 0), recv if recv[1] > recv[0] else (0, 0)))
    return out


def plan_cg_native(plan, vals, b, free=None, rtol=1e-12, atol=0.0, maxiter=None, x0=None, extra=(),
                   scaled_norm=False, check_every=16, graph=None):
    """``pf3_plan_cg``: Jacobi-preconditioned CG for ``(P A P) x = P b`` entirely in native kernels on ONE device,
    ``A = vals + sum(c * v for plan_i, v, c in extra)``.  ``scaled_norm=True`` applies the stopping test
    ``|r| <= max(rtol |b|, atol)`` in the diagonally scaled norm, i.e. exactly what the reference's
    ``cg(D Kuu D, D fu, atol=...)`` measures (tests/test_quad4r_linear_buckling_plate.py:135-146).
    ``graph``: replay each batch of ``check_every`` iterations as one CUDA graph (default: systems below 2 M dofs, which
    are launch-bound); the solve then runs on a side stream ordered after / before the current torch stream.
    Returns (x[6*nnodes], iterations, status, residual norm, |b|); status 0 = converged, 1 = maxiter, 2 = breakdown."""
    dev = vals.device
    n = 6 * plan.nnodes
    if plan.nrows != n:
        raise ValueError("pf3_plan_cg needs a plan that owns every row")
    ctx = context(dev)
    free_t = None if free is None else _dev(free, torch.uint8, dev)
    bt = _dev(b, torch.float64, dev).reshape(-1)
    if bt.numel() != n:
        raise ValueError("b must have 6*nnodes entries")
    x = torch.zeros(n, dtype=torch.float64, device=dev) if x0 is None else _dev(x0, torch.float64, dev).clone()
    plans = [plan._plan] + [pe._plan for pe, _, _ in extra]
    vlist = [_ptr(vals)] + [_ptr(ve) for _, ve, _ in extra]
    coefs = [1.0] + [float(ce) for _, _, ce in extra]
    if graph is None:
        graph = n < 2000000
    flags = (1 if scaled_norm else 0) | (2 if graph else 0) | (int(check_every) << 8)
    args = (plans, vlist, coefs, _ptr(free_t) if free_t is not None else 0, _ptr(bt), _ptr(x), x0 is not None,
            float(rtol), float(atol), int(maxiter or 0), flags)
    if graph:
        # stream capture is refused on the legacy default stream: run on a side stream ordered against the caller's
        cur = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            ctx = context(dev)
            it, status, res, bn = ctx.plan_cg(*args)
        cur.wait_stream(side)
        context(dev)
    else:
        it, status, res, bn = ctx.plan_cg(*args)
    return x, it, status, res, bn


_SIDE = {}


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def compact_csr(indptr, indices, vals, free, row0=0, pattern=None, upper=False, ncols=None, out=None,
                want_indices=True):
    """``K[bu, :][:, bu]`` (tests/test_quad4_static_point_load.py:84-99) as device CSR arrays from a device CSR matrix:
    rows / columns whose ``free`` flag is set are kept and renumbered by their rank among the free dofs.  ``row0``: the
    global index of the matrix' first row (row-sharded plans).  ``free=None`` keeps every dof (``ncols`` is then
    required); ``upper=True`` keeps only ``col >= row`` (``scipy.sparse.triu``: one triangle of a symmetric matrix, 5/9 of
    the values).  Returns ((indptr_uu, indices_uu, vals_uu), pattern); hand ``pattern`` back in to refresh only the
    values of a fixed pattern (``out``: preallocated value tensor)."""
    dev = indptr.device
    ctx = context(dev)
    nrows = indptr.numel() - 1
    free = _dev(free, torch.uint8, dev) if free is not None else None
    if free is not None:
        ncols = free.numel()
    elif ncols is None:
        raise ValueError("ncols is required without a dof mask")
    flags = 1 if upper else 0
    fptr = _ptr(free) if free is not None else 0
    if pattern is None:
        colmap = torch.empty(ncols + 1, dtype=torch.int64, device=dev)
        optr = torch.empty(nrows + 1, dtype=torch.int64, device=dev)
        nkeep, nnz = ctx.csr_compact_symbolic(nrows, ncols, _ptr(indptr), _ptr(indices), fptr, flags, row0,
                                              _ptr(colmap), _ptr(optr))
        optr = optr[:nkeep + 1]
        oidx = torch.empty(nnz, dtype=torch.int64, device=dev) if want_indices else None
        pattern = (colmap, optr, oidx, nnz, not want_indices)
    colmap, optr, oidx, nnz, filled = pattern
    oval = None
    if vals is not None:
        oval = out if out is not None else torch.empty(nnz, dtype=torch.float64, device=dev)
    ctx.csr_compact_fill(nrows, ncols, _ptr(indptr), _ptr(indices), _ptr(vals) if vals is not None else 0, fptr, flags,
                         row0, _ptr(colmap), _ptr(optr), 0 if (filled or oidx is None) else _ptr(oidx),
                         _ptr(oval) if oval is not None else 0)
    return (optr, oidx, oval), (colmap, optr, oidx, nnz, True)


def plan_compact(plan, vals, free, pattern=None, upper=False, out=None, want_indices=True):
    """``compact_csr`` on the CSR values of an ``AssemblyPlan`` (its own row block when row-sharded)."""
    indptr, indices = plan.pattern()
    return compact_csr(indptr, indices, vals, free, row0=6 * plan.node_begin, pattern=pattern, upper=upper,
                       ncols=6 * plan.nnodes, out=out, want_indices=want_indices)


def shift_invert_operator(plan_a, a_vals, plan_m, m_vals, sigma, free, rtol=1e-13, maxiter=None):
    """``OPinv`` for ``scipy.sparse.linalg.eigsh(A=Kuu, M=Muu, sigma=sigma, OPinv=...)``: x -> (A - sigma M)^-1 x on
    the free DOFs, each application one Jacobi-CG on the device with the operator applied as two block SpMVs
    (the natural-frequency scripts call ``eigsh(A=Kuu, M=Muu, sigma=-1.)``,
    tests/test_beamc_natural_freq_curved.py:107, tests/test_tria3r_natural_freq_distorted.py:150).
    ``A - sigma M`` must be positive definite (sigma below the spectrum, e.g. the scripts' -1).  Vectors are indexed
    like ``K[bu, :][:, bu]``: by the free DOFs in order."""
    import numpy as np
    from scipy.sparse.linalg import LinearOperator
    dev = a_vals.device
    free_t = _dev(free, torch.uint8, dev)
    idx = torch.nonzero(free_t).ravel()
    nfree = int(idx.numel())
    full = torch.zeros(free_t.numel(), dtype=torch.float64, device=dev)
    stats = {"solves": 0, "iterations": 0}

    def matvec(v):
        full.zero_()
        full[idx] = torch.as_tensor(np.asarray(v, float).ravel()).to(dev)
        x, info = plan_cg_solve(plan_a, a_vals, full, free=free_t, rtol=rtol, maxiter=maxiter,
                                extra=[(plan_m, m_vals, -float(sigma))])
        if info < 0:
            raise RuntimeError("shift-invert CG did not converge")
        stats["solves"] += 1
        stats["iterations"] += info
        return x[idx].cpu().numpy()

    op = LinearOperator((nfree, nfree), matvec=matvec, dtype=np.float64)
    op.stats = stats
    return op
