"""GPU, world_size 2, NCCL (skipped with fewer than two GPUs): config 1 evaluated, assembled and solved with the rows
sharded over two ranks -- each rank runs the fused kernel on the elements touching its node range, keeps its own
CSR row block, and the Jacobi-CG exchanges only the halo of the search direction (grouped NCCL send / recv per iteration)."""
import json
import os
import socket

import numpy as np
import pytest

from tests import configs, util

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from pyfe3d_b200 import sharding
    from pyfe3d_b200.batch import AssemblyPlan
    from pyfe3d_b200.solve import plan_cg_solve
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    c = configs.build("quad4_static")[0]
    n = c["ndof"]
    sub = sharding.shard_case(c, rank, world)
    b = util.batch_from_case(sub, device=torch.device("cuda", rank))
    plan = AssemblyPlan("KC0", n // 6, [b], node_range=sub["owned_nodes"])
    _, csr = plan.evaluate_assemble(KC0=True, write_coo=False)
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
    dev = torch.device("cuda", rank)
    u, info = plan_cg_solve(plan, csr["KC0"], torch.as_tensor(f).to(dev),
                            free=torch.as_tensor((~bk).astype(np.uint8)).to(dev), rtol=1e-13, group=dist.group.WORLD)
    q.put((rank, info, float(u[2::6].max()), plan.nrows))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_static_solve_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gold = json.load(open(os.path.join(util.GOLDEN_DIR, "config_scalars.json")))
    want = gold["quad4_static"]["w_max"]
    c = configs.build("quad4_static")[0]
    assert sum(t[3] for t in res) == c["ndof"]
    for _, info, w_max, _ in res:
        assert info > 0
        assert abs(w_max - want) <= 1e-8 * abs(want)
