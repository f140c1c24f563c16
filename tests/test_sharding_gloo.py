"""CPU, world_size 2, gloo: the host-side row-ownership logic of the multi-GPU path
(node ranges, halo element selection, nnz offset gather).  The device work itself is covered by
tests/test_gpu_fused.py::test_fused_given_stress_and_shard and the lumped-mass shard test."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import coo as ocoo
from oracle import driver
from pyfe3d_b200 import sharding
from tests import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = cases.shell_mesh("quad4", 9, 8, seed=3)
    n = case["ndof"]
    sub = sharding.shard_case(case, rank, world)
    lo, hi = sub["owned_nodes"]
    # each rank assembles ONLY the rows it owns, from the elements touching them (oracle stands in for
    # the device kernels here: this test is about the partition logic)
    out = driver.run(sub, what=("KC0",))
    r, c, v = out["KC0"]
    keep = (r >= 6 * lo) & (r < 6 * hi)
    A = sp.coo_matrix((v[keep], (r[keep] - 6 * lo, c[keep])), shape=(6 * (hi - lo), n)).tocsr()
    A.sum_duplicates()
    off, sizes = sharding.gather_row_blocks(torch.zeros(1), A.nnz)
    q.put((rank, lo, hi, A, off, sizes, len(sub["element_ids"]), len(sharding.owned_elements(case["conn"], lo, hi))))
    dist.barrier()
    dist.destroy_process_group()


def test_row_ownership_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = cases.shell_mesh("quad4", 9, 8, seed=3)
    n = case["ndof"]
    full = driver.run(case, what=("KC0",))["KC0"]
    S = sp.coo_matrix((full[2], (full[0], full[1])), shape=(n, n)).tocsr()
    S.sum_duplicates()
    stacked = sp.vstack([t[3] for t in res]).tocsr()
    assert abs(stacked - S).max() <= 1e-11 * abs(S).max()
    assert stacked.nnz == S.nnz
    # nnz offsets from the gather: rank 1 starts where rank 0 ends
    assert res[0][4] == 0 and res[1][4] == res[0][3].nnz and res[0][5] == res[1][5]
    # halo: evaluated elements > owned elements, owned elements form a disjoint cover
    assert sum(t[7] for t in res) == case["conn"].shape[0]
    assert all(t[6] >= t[7] for t in res) and sum(t[6] for t in res) > case["conn"].shape[0]


def test_node_ranges_cover():
    for nn, w in ((10, 3), (7, 8), (4001 * 2001, 8)):
        rs = sharding.node_ranges(nn, w)
        assert rs[0][0] == 0 and rs[-1][1] == nn
        assert all(a[1] == b[0] for a, b in zip(rs[:-1], rs[1:]))


def _halo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from pyfe3d_b200.solve import halo_exchange, halo_plan
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = cases.shell_mesh("quad4", 9, 8, seed=3)
    n = case["ndof"]
    sub = sharding.shard_case(case, rank, world)
    lo, hi = sub["owned_nodes"]
    conn = np.asarray(sub["conn"])
    mine = torch.tensor([6 * lo, 6 * hi, min(6 * int(conn.min()), 6 * lo), max(6 * int(conn.max()) + 6, 6 * hi)])
    table = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(table, mine)
    table = [t.tolist() for t in table]
    halo = halo_plan(table, rank)
    ref = torch.arange(n, dtype=torch.float64) * 0.5 + 1.0        # the "global" vector every rank should see
    vec = torch.full((n,), -7.0, dtype=torch.float64)
    vec[6 * lo:6 * hi] = ref[6 * lo:6 * hi]                         # only the own rows are current
    halo_exchange(vec, halo, dist.group.WORLD)
    need = slice(table[rank][2], table[rank][3])
    cols = np.unique(conn)                                          # nodes whose columns this rank's rows read
    dofs = (6 * cols[:, None] + np.arange(6)).ravel()
    q.put((rank, bool(torch.equal(vec[need], ref[need])), bool(torch.equal(vec[dofs], ref[dofs])),
           float((vec == -7.0).sum()), halo))
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_two_ranks_gloo():
    """The point-to-point halo exchange of the sharded CG (pyfe3d_b200.solve.halo_plan / halo_exchange): after one
    exchange every rank holds the reference values at every column its rows read, and nothing else moved."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, range_ok, cols_ok, untouched, halo in res:
        assert range_ok and cols_ok
        assert untouched > 0                       # the far part of the vector was not transferred
        assert len(halo) == 1 and halo[0][0] == 1 - rank


def test_halo_plan_banded_three_ranks():
    from pyfe3d_b200.solve import halo_plan
    table = [(0, 60, 0, 72), (60, 120, 48, 132), (120, 180, 108, 180)]
    assert halo_plan(table, 0) == [(1, (48, 60), (60, 72))]
    assert halo_plan(table, 1) == [(0, (60, 72), (48, 60)), (2, (108, 120), (120, 132))]
    assert halo_plan(table, 2) == [(1, (120, 132), (108, 120))]
    # every send has the matching receive on the other side
    for me in range(3):
        for r, send, recv in halo_plan(table, me):
            back = [h for h in halo_plan(table, r) if h[0] == me][0]
            assert back[2] == send and back[1] == recv
    # a rank whose rows read nothing outside them exchanges nothing
    assert halo_plan([(0, 60, 0, 60), (60, 120, 60, 120)], 0) == []
