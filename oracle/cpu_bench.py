"""CPU arm of bench.py: the reference's own element loop + scipy assembly, on the host cores.

TEST/BENCH INFRASTRUCTURE ONLY.  Runs the UNMODIFIED compiled reference
(``oracle/_ref``, built from /root/reference by ``oracle/build_ref.py``) exactly the way
its tests drive it: per element ``update_rotation_matrix``, ``update_probe_xe``,
``update_probe_ue``, ``update_KC0``, ``update_KG``, ``update_M`` into COO arrays
(tests/test_quad4_static_point_load.py:53-78, tests/test_quad4r_linear_buckling_plate.py:163-166),
then ``coo_matrix(...).tocsr()`` (:80).  The reference has no parallel mode of its own; to use
all host cores the element range is cut into one contiguous slice per process, each with its
own probe and COO arrays (SURVEY §8(d) "CPU baseline beside it").
"""
import multiprocessing as mp
import os
import time

import numpy as np

_STATE = {}


def _init(side):
    from oracle import ref_loop
    from pyfe3d_b200 import meshes
    ref_loop.load()
    _STATE["case"] = meshes.plate_quad4(side, side)
    _STATE["ref_loop"] = ref_loop


def _work(rng):
    import scipy.sparse as sp
    e0, e1 = rng
    case = _STATE["case"]
    t0 = time.perf_counter()
    out = _STATE["ref_loop"].run(case, what=("KC0", "KG", "M0"), e0=e0, e1=e1)
    t1 = time.perf_counter()
    n = case["ndof"]
    nnz = 0
    for k in ("KC0", "KG", "M0"):
        r, c, v = out[k]
        nnz += sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr().nnz
    return nnz, t1 - t0, time.perf_counter() - t1


class ReferenceBench:
    """Persistent worker pool so that imports and mesh generation stay outside the timed step."""

    def __init__(self, side=256, nproc=None):
        self.side = side
        self.nproc = nproc or os.cpu_count() or 1
        self.ne = side * side
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.nproc, initializer=_init, initargs=(side,))
        cuts = np.linspace(0, self.ne, self.nproc + 1).astype(int)
        self.slices = [(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        self.pool.map(_work, [(0, 8)] * self.nproc)   # touch every worker once (untimed)

    def step(self):
        t0 = time.perf_counter()
        res = self.pool.map(_work, self.slices, chunksize=1)
        dt = time.perf_counter() - t0
        # slowest worker's split: element loop against scipy coo -> csr
        self.loop_s = max(r[1] for r in res)
        self.tocsr_s = max(r[2] for r in res)
        return dt

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self):
        return ("%dx%d-element sub-plate of the workload mesh (%d Quad4), reference Cython loop "
                "rot+xe+ue+KC0+KG+M(mtype0) with indices, then scipy coo->csr per slice; %d processes"
                % (self.side, self.side, self.ne, self.nproc))


def reference_static_solve(side, rtol=1e-9):
    """Reference arm of the static solve on the host: Cython element loop (KC0) -> coo_matrix.tocsc ->
    K[bu,:][:,bu] -> diagonally scaled scipy cg.  Returns (seconds, iterations, max normal displacement)."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import cg
    from oracle import ref_loop
    ref_loop.load()
    from pyfe3d_b200 import meshes
    case, free, f, normal = meshes.static_case(side)
    n = case["ndof"]
    t0 = time.perf_counter()
    out = ref_loop.run(case, what=("KC0",))
    r, c, v = out["KC0"]
    K = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsc()
    Kuu = K[free, :][:, free]
    dis = 1.0 / np.sqrt(np.maximum(Kuu.diagonal(), 1e-30))
    D = sp.diags(dis)
    fs = D @ f[free]
    its = [0]

    def cb(_):
        its[0] += 1
    us, info = cg(D @ Kuu @ D, fs, rtol=0., atol=rtol * np.linalg.norm(fs), maxiter=200000, callback=cb)
    u = np.zeros(n)
    u[free] = D @ us
    dt = time.perf_counter() - t0
    w = u[0::6] * normal[0] + u[1::6] * normal[1] + u[2::6] * normal[2]
    return dt, its[0], float(np.abs(w).max()), info
