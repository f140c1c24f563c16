#!/usr/bin/env python
"""BASELINE config 5: stiffened panel, Quad4 skin 3968 x 3968 (15.75 M) + BeamC stiffeners along 64 grid lines (0.254 M),
KC0 / KG / M in ONE matrix each + update_fint, sharded over N GPUs by DOF-row ownership (weak scaling: every rank owns a
strip of 496 node columns with its 8 stiffener lines; halo elements duplicated; no collective on the data path).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_config5.py [--cols 496 --rows 3968 --steps 5]
Prints one JSON line (rank 0).  Not the headline metric (bench.py is)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfe3d_b200 import meshes  # noqa: E402
from pyfe3d_b200.batch import AssemblyPlan, Coo, ElementBatch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cols", type=int, default=496)
    ap.add_argument("--rows", type=int, default=3968)
    ap.add_argument("--lines", type=int, default=8, help="stiffener lines per rank")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nx, ny = a.cols * world, a.rows
    i0 = rank * a.cols + (1 if rank > 0 else 0)
    i1 = (rank + 1) * a.cols + 1
    skin = meshes.plate_quad4(nx, ny, a=float(world) * a.cols / a.rows, b=1.0, i0=i0 if world > 1 else None,
                              i1=i1 if world > 1 else None, local=True)
    nn = skin["ndof"] // 6
    nny = ny + 1
    lo, hi = skin["owned_nodes"]
    c_lo, c_hi = lo // nny, hi // nny                       # owned node columns (local numbering)
    lines = np.linspace(c_lo, c_hi, a.lines + 2)[1:-1].round().astype(int)
    n1 = (np.repeat(lines, ny) * nny + np.tile(np.arange(ny), lines.size)).astype(np.int64)
    bconn = np.stack([n1, n1 + 1], 1)
    E, nu, rho, bb, hh = 70e9, 0.33, 2700., 0.002, 0.02
    A, Iyy, Izz = bb * hh, bb * hh ** 3 / 12, hh * bb ** 3 / 12
    p = np.zeros((1, 16))
    p[0, :9] = [A, E, E / 2 / (1 + nu) * 5 / 6., Iyy, Izz, 0., Iyy + Izz, 0., 0.]
    p[0, 9:15] = [rho * A, 0., 0., rho * Izz, rho * Iyy, 0.]
    normal = meshes.fixed_rotation(0)[:, 2]
    bs = [ElementBatch("quad4", skin["conn"], skin["x"], skin["props"], u=skin["u"], device=dev),
          ElementBatch("beamc", bconn, skin["x"], p, u=skin["u"], vxy=np.tile(normal, (bconn.shape[0], 1)), nnodes=nn,
                       device=dev)]
    plan = AssemblyPlan("KC0", nn, bs, node_range=(lo, hi))
    names = ("KC0", "KG", "M")
    plans = {m: plan._sibling(m, 0) for m in names}
    coo = {m: Coo(None, None, torch.zeros(plans[m].coo_size, dtype=torch.float64, device=dev), 6 * nn) for m in names}
    csr = {m: torch.empty(plans[m].nnz, dtype=torch.float64, device=dev) for m in names}
    fint = torch.zeros(6 * nn, dtype=torch.float64, device=dev)

    def step():
        plan.evaluate_assemble(KC0=True, KG=True, M=True, coo=coo, csr=csr)
        plan.update_fint(fint)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        step()
    e.record()
    barrier()
    t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / a.steps
    unique = nx * ny + a.lines * world * ny
    if rank == 0:
        print(json.dumps({"config": "config5 stiffened panel: %d x %d Quad4 + %d BeamC, KC0+KG+M (one matrix each) + fint"
                                    % (nx, ny, a.lines * world * ny), "n_gpus": world, "elements_total": unique,
                          "elements_evaluated_per_gpu": int(bs[0].ne + bs[1].ne), "ms_per_step": ms,
                          "elements_per_s": unique / ms * 1e3, "fused_quad_share": not getattr(plan, "_fused_unsupported", False),
                          "nnz_per_gpu": {m: int(plans[m].nnz) for m in names}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
