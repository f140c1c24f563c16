#!/usr/bin/env python
"""Row-sharded Jacobi-CG over N GPUs (SURVEY 8(e) + 8(f) rank 1): KC0 of the north-star plate assembled by the fused kernel
on every rank's row block, then `plan_cg_solve(group=WORLD)`: block SpMV on the own rows, point-to-point halo exchange of
the search direction, two small all_reduces per iteration.  Prints ms per iteration (slope between two iteration counts,
device-timed, max over ranks) for the halo exchange and for the all_gather of the whole direction it replaces.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_cg_multi.py [side]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfe3d_b200 import meshes, sharding  # noqa: E402
from pyfe3d_b200.batch import AssemblyPlan  # noqa: E402
from pyfe3d_b200.solve import plan_cg_solve  # noqa: E402


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD if world > 1 else None

    def setup(sd):
        case, free, f, normal = meshes.static_case(sd)
        n = case["ndof"]
        sub = sharding.shard_case(case, rank, world)
        b = meshes.batch_from_case(sub, device=dev)
        plan = AssemblyPlan("KC0", n // 6, [b], node_range=sub["owned_nodes"])
        _, csr = plan.evaluate_assemble(KC0=True, write_coo=False)
        return plan, csr["KC0"], torch.as_tensor(free.astype(np.uint8)).to(dev), torch.as_tensor(f).to(dev), normal, n

    plan, vals, free_t, ft, normal, n = setup(side)

    def timed(iters, graph):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x, info = plan_cg_solve(plan, vals, ft, free=free_t, rtol=0., maxiter=iters, group=group, graph=graph)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t), x

    out = {"config": "row-sharded Jacobi-CG, KC0 of %dx%d Quad4 (%d dofs), clamped edges" % (side, side, n), "n_gpus": world,
           "rows_per_rank": plan.nrows, "nnz_per_rank": plan.nnz}
    k1, k2 = 64, 320
    modes = (("halo", False), ("all_gather", False), ("halo", True)) if world > 1 else (("halo", False),)
    for mode, graph in modes:
        os.environ["PF3_CG_EXCHANGE"] = mode
        timed(32, graph)                              # warm-up (NCCL channels, allocator)
        t1, _ = timed(k1, graph)
        t2, x = timed(k2, graph)
        out["ms_per_iteration_%s_%s" % (mode, "graph" if graph else "eager")] = (t2 - t1) / (k2 - k1)
    os.environ["PF3_CG_EXCHANGE"] = "halo"
    del plan, vals
    # correctness bit on a mesh this Jacobi-CG converges on: residual of the sharded operator at the returned solution
    plan, vals, free_t, ft, normal, n = setup(64)
    xs, info = plan_cg_solve(plan, vals, ft, free=free_t, rtol=1e-10, maxiter=200000, group=group)
    lo, hi = 6 * plan.node_begin, 6 * plan.node_end
    bl = (ft * free_t)[lo:hi]
    r = bl - plan.spmv(vals, xs, free=free_t)
    rr = torch.stack([torch.dot(r, r), torch.dot(bl, bl)])
    dist.all_reduce(rr)
    out["check_64x64"] = {"iterations": info, "relative_residual": float((rr[0] / rr[1]).sqrt()),
                          "max_deflection": float((xs.view(-1, 6)[:, :3] @ torch.as_tensor(normal).to(dev)).abs().max())}
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
