#!/usr/bin/env python
"""bench.py — Quad4 KC0+KG+M evaluation + CSR assembly throughput (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # reference Cython loop on host cores

One step = one pass of the hot path over the whole mesh: the record kernel + ONE fused node-centric kernel writing
the KC0/KG/M COO value arrays (values only — the steady-state `update_*v_only` case; the index arrays are a function
of connectivity and are not re-written) AND the three assembled CSR value arrays through a precomputed plan.
N>1 (torchrun): node rows are strip-partitioned, each rank evaluates the elements touching its rows (halo duplicated)
and assembles its own CSR row block: no collective on the data path.  The headline `value` is WEAK scaling (the mesh
grows to N x side x side elements); the same line carries, under `details`,
  * `strong`  : the metric's own size (side x side = 4.0 M elements in total) cut into N strips,
  * `others`  : BASELINE configs 2-5 timed with the same clock (N=1: config 5 as one GPU's share; N>1: sharded),
  * `solve_e2e`: a user-level end-to-end — host x in, KC0 assembled and the static problem solved by the native CG on the
                device, only u back — next to the reference doing loop + tocsc + scipy cg on the same mesh,
  * `sharded_cg`: (N>1) the one downstream exchange step — Jacobi-CG on KC0 with the rows sharded over the ranks: block SpMV
                on the own rows + point-to-point halo exchange of the search direction (grouped NCCL send / recv), ms per
                iteration, with the all_gather it replaces beside it and the true residual as a correctness bit,
and a `parity` object: checksums that tie every rank's CSR block to its COO arrays and to closed forms, all-reduced, so
that every line of a scaling run carries a correctness bit.
Prints one JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "quad4_kc0_kg_m_eval_plus_csr_assembly_throughput"
UNIT = "elements/s"
# SURVEY §8(d) algorithmic bytes per Quad4 element (values-only steady state)
BYTES_EVAL = (576 + 144 + 480) * 8 + 108            # 9600 B of COO values + connectivity/coords/u reads
BYTES_CSR = (324 + 81 + 270) * 8                    # 5400 B of CSR values (amortised per element)
BYTES_ASM_READ = (576 + 144 + 480) * 8              # assembly re-reads the COO values
BYTES_PATH = BYTES_EVAL + BYTES_CSR                 # 15.1 kB: fused lower bound of the whole step


def clocks_sampler(index, stop, out):
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                              "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return

    def reader():
        for line in p.stdout:
            out.append(line.strip())
    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop.wait()
    p.terminate()
    t.join(timeout=2)


def summarize_clocks(lines):
    sm, mx, reasons = [], 0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [t.strip() for t in ln.split(",")]
        if len(f) < 6:
            continue
        try:
            sm.append(float(f[0]))
            mx = max(mx, float(f[1]))
        except ValueError:
            continue
        for n, v in zip(names, f[2:6]):
            if v.lower().startswith("active"):
                reasons.add(n)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


L2_NOTE = ("every timed step writes >= 15 kB per element of fresh COO + CSR output (60 GB at 4.0 M elements), far more than "
           "the 126 MB L2: inputs and outputs never stay cached between steps, no flush needed")


def workload_name(side):
    """config.workload, shared by both arms (the reference arm times a bounded sample of it)."""
    return ("configs[1..] north-star mesh: %dx%d Quad4 structured plate per GPU (%d elements/GPU), rigidly rotated, "
            "coupled laminate [30,-45,0] offset 0.5 mm, u=1e-4 N(0,1); KC0+KG(from u)+M(mtype 0) values + CSR assembly"
            % (side, side, side * side))


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref_loop
    if not ref_loop.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (python oracle/build_ref.py)"}))
        return 0
    from oracle.cpu_bench import ReferenceBench
    rb = ReferenceBench(side=args.cpu_side)
    for _ in range(args.warmup):
        rb.step()
    t = [rb.step() for _ in range(args.steps)]
    rb.close()
    total = sum(t)
    value = rb.ne * args.steps / total
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.side), "l2": L2_NOTE},
            "details": {"elements_per_step": rb.ne,
                        "note": "each step is a bounded sample of that workload (a %dx%d-element sub-plate of the same "
                                "mesh, laminate and displacement field); the reference cannot hold 4M Quad4 in one COO "
                                "array (int32 init_k, quad4.pyx:453)" % (args.cpu_side, args.cpu_side)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": rb.nproc, "kind": "reference", "sample": rb.describe()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ our arm
class Dist:
    """torch.distributed plumbing of the bench (NCCL: barrier, max / sum reductions of a few scalars)."""

    def __init__(self, torch, dist, dev, world):
        self.torch, self.dist, self.dev, self.world = torch, dist, dev, world

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op="sum"):
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX,
                                        "min": self.dist.ReduceOp.MIN}[op])
        return [float(v) for v in t.tolist()]

    def gather(self, val):
        t = self.torch.tensor([float(val)], dtype=self.torch.float64, device=self.dev)
        if self.world == 1:
            return [float(val)]
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]


def plate_shard(meshes, nx, ny, a, rank, world):
    """Rank `rank`'s strip of the nx x ny plate: it owns nx/world element columns' worth of node columns."""
    cols = nx // world
    i0 = rank * cols + (1 if rank > 0 else 0)
    i1 = (rank + 1) * cols + 1 if rank < world - 1 else nx + 1
    return meshes.plate_quad4(nx, ny, a=a, b=1.0, i0=i0 if world > 1 else None, i1=i1 if world > 1 else None,
                              local=True)


def parity_checks(torch, D, case, batch, plans, coos, csr, a_len, ne_total):
    """Correctness bits of this run, all-reduced over the ranks.
      * csr_vs_coo[m]: |sum(CSR values of the owned rows) - sum(COO row slabs of the owned nodes)| / sum|CSR|.  A rank
        writes the COO slab of an (element, local node) pair exactly when it owns that node's rows, so over all ranks
        both are the sum of every entry of the global matrix: a wrong halo, row cut or slot map breaks the equality.
      * mass: sum(M e_x) over the x-translation dofs against intrho * a * b (closed form for the plate).
      * rigid: max|KC0 e_x| / max|KC0| (a rigid translation produces no force).
      * kc0_abs_per_element / m_abs_per_element (sum of |entries| / elements): every element of this mesh has the
        same matrices, so these do not depend on N: compare them between the lines of a scaling run (KG depends on u
        and is tied by csr_vs_coo)."""
    lo, hi = case["owned_nodes"]
    conn0 = batch.conn[:, 0]
    owned = (conn0 >= lo) & (conn0 < hi)                 # unique cover of the elements (first node owned)
    own_slab = (batch.conn >= lo) & (batch.conn < hi)    # (element, local node) row slabs this rank writes
    ne = batch.ne
    out, loc = {}, []
    for m in ("KC0", "KG", "M"):
        v = coos[m].v.view(ne, batch.nn, -1)             # row slab of local node a = entries [a, a+1) * size / nn
        loc += [float(csr[m].sum()), float(csr[m].abs().sum()), float(v[own_slab].sum()),
                float(v[own_slab].abs().sum())]
    loc.append(float(owned.sum()))
    tot = D.reduce(loc)
    worst = 0.
    for i, m in enumerate(("KC0", "KG", "M")):
        s_csr, s_abs, s_coo = tot[4 * i:4 * i + 3]
        rel = abs(s_csr - s_coo) / s_abs if s_abs > 0 else float("inf")
        out["csr_vs_coo_" + m] = rel
        worst = max(worst, rel)
    ne_unique = tot[12]
    out["unique_elements"] = int(ne_unique)
    out["kc0_abs_per_element"] = tot[3] / ne_unique
    out["m_abs_per_element"] = tot[11] / ne_unique
    nn = case["ndof"] // 6
    ex = torch.zeros(6 * nn, dtype=torch.float64, device=csr["M"].device)
    ex[0::6] = 1.0
    ym = plans["M"].spmv(csr["M"], ex)
    yk = plans["KC0"].spmv(csr["KC0"], ex)
    mass = D.reduce([float(ym[0::6].sum())])[0]
    kmax = D.reduce([float(yk.abs().max()), float(csr["KC0"].abs().max())], "max")
    intrho = float(case["props"][0, 24])
    out["mass_rel_err"] = abs(mass - intrho * a_len * 1.0) / (intrho * a_len)
    out["rigid_translation_residual"] = kmax[0] / kmax[1]
    out["ok"] = bool(worst <= 1e-11 and out["mass_rel_err"] <= 1e-10 and out["rigid_translation_residual"] <= 1e-10
                     and int(ne_unique) == int(ne_total))
    return out


def measure_plate(torch, D, meshes, nx, ny, a_len, rank, world, dev, args, full):
    """Plan + warm-up + timed steps of the fused path on this rank's strip of the nx x ny plate.  `full`: also the
    host-buffer end-to-end steps.  Returns a dict (times are max over ranks)."""
    import numpy as np
    from pyfe3d_b200.batch import AssemblyPlan, ElementBatch, context
    case = plate_shard(meshes, nx, ny, a_len, rank, world)
    nnodes = case["ndof"] // 6
    ne_local = case["conn"].shape[0]
    batch = ElementBatch("quad4", case["conn"], case["x"], case["props"], u=case["u"], device=dev)
    ctx = context(dev)
    mats = ("KC0", "KG", "M")
    t_sym0 = time.perf_counter()
    plans = {m: AssemblyPlan(m, nnodes, [batch], node_range=case["owned_nodes"]) for m in mats}
    torch.cuda.synchronize()
    t_symbolic = time.perf_counter() - t_sym0
    coos = batch.evaluate(KC0=True, KG=True, M=True, indices=False)          # allocates the value arrays
    csr = {m: torch.empty(plans[m].nnz, dtype=torch.float64, device=dev) for m in mats}
    fused = args.path == "fused"

    def step(ev=None):
        if ev is not None:
            ev[0].record()
        if fused:
            plans["KC0"].evaluate_assemble(KC0=True, KG=True, M=True, coo=coos, csr=csr)
            if ev is not None:
                for i in range(1, 5):
                    ev[i].record()
            return
        batch.evaluate(KC0=True, KG=True, M=True, indices=False, out=coos)
        if ev is not None:
            ev[1].record()
        for i, m in enumerate(mats):
            plans[m].assemble(coos[m].v, out=csr[m])
            if ev is not None:
                ev[2 + i].record()

    for _ in range(max(args.warmup, 3)):
        step()
    D.barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    clk_lines, stop = [], threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(dev.index, stop, clk_lines), daemon=True)
    th.start()
    time.sleep(0.25)
    l0 = ctx.launch_count()
    D.barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for k in range(args.steps):
        step(evs[k])
    end.record()
    D.barrier()
    launches = ctx.launch_count() - l0
    stop.set()
    th.join(timeout=3)
    kern = np.zeros(4)
    for ev in evs:
        for i in range(4):
            kern[i] += ev[i].elapsed_time(ev[i + 1])
    kern /= args.steps                                         # ms per launch: eval, asm KC0, asm KG, asm M
    ms_step = D.reduce([start.elapsed_time(end)], "max")[0] / args.steps
    res = {"ms_step": ms_step, "kern": kern, "ne_local": ne_local, "ne_total": nx * ny, "launches": int(launches),
           "symbolic_s": t_symbolic, "clocks": summarize_clocks(clk_lines), "fused": fused}
    res["parity"] = parity_checks(torch, D, case, batch, plans, coos, csr, a_len, nx * ny)

    # ---- end to end through the public API with host buffers --------------------------------------
    if full and args.e2e_steps > 0:
        try:
            # page-locked result buffers: portable (every context of the process may use them) and write-combined (the
            # GPU's writes do not snoop the CPU caches).  numa_local first-touches them from the GPU's own NUMA node.
            with numa_local(dev.index) as numa:
                xh = torch.as_tensor(case["x"]).pin_memory()
                uh = torch.as_tensor(case["u"]).pin_memory()
                outh = {m: torch.empty(plans[m].nnz, dtype=torch.float64).pin_memory() for m in mats}
                for o in outh.values():
                    o.zero_()
            h2d = xh.numel() * 8 + uh.numel() * 8
            d2h = sum(o.numel() * 8 for o in outh.values())

            def e2e_step():
                if fused:
                    # ONE C-ABI call with host buffers (pf3_eval_assemble_host): H2D of x, u -> K1 + K2 -> D2H of
                    # the three CSR value arrays; returns after the copies have completed
                    plans["KC0"].evaluate_assemble_host(xh, uh, outh, KC0=True, KG=True, M=True, coo=coos)
                    return
                batch.x.copy_(xh, non_blocking=True)
                batch.u.copy_(uh, non_blocking=True)
                step()
                for m in mats:
                    outh[m].copy_(csr[m], non_blocking=True)
                torch.cuda.synchronize()
            e2e_step()
            D.barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            t_own = time.perf_counter() - t0
            D.barrier()
            dt = D.reduce([time.perf_counter() - t0], "max")[0]
            per_rank = D.gather(d2h * args.e2e_steps / t_own / 1e9)
            # the box's raw device->host rate for the same buffers (one plain copy of the KC0 values, nothing else
            # running): what the end-to-end step is limited by
            torch.cuda.synchronize()
            t0r = time.perf_counter()
            outh["KC0"].copy_(csr["KC0"], non_blocking=True)
            torch.cuda.synchronize()
            raw_gbs = outh["KC0"].numel() * 8 / (time.perf_counter() - t0r) / 1e9
            res["e2e"] = {"value": nx * ny * args.e2e_steps / dt, "unit": UNIT,
                          "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "numa": numa,
                          "d2h_gbs_per_rank": [round(v, 2) for v in per_rank],
                          "d2h_gbs_raw_copy_rank0": round(raw_gbs, 2),
                          "what": "one pf3_eval_assemble_host call per step (C ABI, host buffers): H2D of x,u from pinned "
                                  "host memory -> record + fused kernels -> D2H of the KC0/KG/M CSR value arrays into "
                                  "pinned host memory (pattern is static; the COO value arrays are written on the device "
                                  "as in the timed steps); d2h_gbs_per_rank names the limiter: the call is the PCIe copy"}
            # ---- the same step returning ONE TRIANGLE of each (symmetric) matrix: 5/9 of the bytes over PCIe.  Same
            # C ABI: H2D of x, u -> pf3_eval_assemble (device) -> pf3_csr_compact_fill(PF3_COMPACT_UPPER) per matrix
            # -> D2H of the compacted values.  Reported beside the headline, not instead of it: the default result of
            # the path is the full scipy-shaped CSR matrix.
            if fused and args.e2e_upper:
                from pyfe3d_b200.solve import plan_compact
                pats, up, uph = {}, {}, {}
                for m in mats:
                    (_, _, v), pats[m] = plan_compact(plans[m], csr[m], None, upper=True, want_indices=False)
                    up[m] = v
                for m in mats:
                    uph[m] = outh[m][:up[m].numel()]          # the pinned buffers of the full-matrix step, reused
                d2h_up = sum(o.numel() * 8 for o in uph.values())

                def upper_step():
                    batch.x.copy_(xh, non_blocking=True)
                    batch.u.copy_(uh, non_blocking=True)
                    step()
                    for m in mats:
                        plan_compact(plans[m], csr[m], None, pattern=pats[m], upper=True, out=up[m])
                        uph[m].copy_(up[m], non_blocking=True)
                    torch.cuda.synchronize()
                upper_step()
                D.barrier()
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    upper_step()
                D.barrier()
                dtu = D.reduce([time.perf_counter() - t0], "max")[0]
                res["e2e"]["symmetric_upper"] = {
                    "value": nx * ny * args.e2e_steps / dtu, "unit": UNIT, "d2h_bytes_per_step": int(d2h_up),
                    "what": "same step, only entries with col >= row of KC0/KG/M returned (scipy.sparse.triu layout, "
                            "PF3_COMPACT_UPPER); for consumers that take a symmetric half"}
                del up, uph, pats
            del outh, xh, uh
        except RuntimeError as exc:   # e.g. pinned allocation refused
            res["e2e"] = {"value": None, "unit": UNIT, "error": str(exc)[:200]}
    del plans, coos, csr, batch
    torch.cuda.empty_cache()
    return res


def time_steps(torch, D, fn, steps, warm=3):
    for _ in range(warm):
        fn()
    D.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    D.barrier()
    return D.reduce([a.elapsed_time(b)], "max")[0] / steps


def other_config(torch, D, meshes, name, case, mats, peak, steps):
    """BASELINE configs 2-4 on one GPU: evaluate + assemble (fused where the kind has it), values only."""
    from pyfe3d_b200.batch import AssemblyPlan
    b = meshes.batch_from_case(case)
    nn = case["ndof"] // 6
    kw = {}
    for m in mats:
        if m == "KGs":
            kw["KG_given_stress"] = case.get("stress", (0., 0., 1.))
        elif m.startswith("M"):
            kw["M"] = True
            kw["mtype"] = int(m[1])
        else:
            kw[m] = True
    plan = AssemblyPlan("KC0", nn, [b])
    coo, csr = plan.evaluate_assemble(**kw)
    ms = time_steps(torch, D, lambda: plan.evaluate_assemble(coo=coo, csr=csr, **kw), steps)
    alg = sum(coo[m].v.numel() * 8 for m in coo) + sum(csr[m].numel() * 8 for m in csr)
    out = {"config": name, "kind": case["kind"], "elements": int(b.ne), "matrices": list(mats),
           "ms_per_step": ms, "elements_per_s": b.ne / ms * 1e3, "algorithmic_bytes": int(alg),
           "achieved_gbs": alg / ms / 1e6, "frac": alg / ms / 1e6 / peak,
           "path": "two-pass" if getattr(plan, "_fused_unsupported", False) or case["kind"] not in ("quad4", "quad4r", "tria3r") else "fused"}
    del plan, coo, csr, b
    torch.cuda.empty_cache()
    return out


def config5(torch, D, meshes, rank, world, dev, cols, rows, lines, peak, steps):
    """BASELINE config 5: stiffened panel, Quad4 skin + BeamC stiffeners in ONE matrix each (KC0, KG, M) + update_fint,
    strip-sharded by DOF-row ownership.  N ranks x `cols` node columns x `rows`; `lines` stiffener lines per rank."""
    import numpy as np
    from pyfe3d_b200.batch import AssemblyPlan, Coo, ElementBatch
    nx, ny = cols * world, rows
    i0 = rank * cols + (1 if rank > 0 else 0)
    i1 = (rank + 1) * cols + 1
    skin = meshes.plate_quad4(nx, ny, a=float(world) * cols / rows, b=1.0, i0=i0 if world > 1 else None,
                              i1=i1 if world > 1 else None, local=True)
    nn = skin["ndof"] // 6
    nny = ny + 1
    lo, hi = skin["owned_nodes"]
    c_lo, c_hi = lo // nny, hi // nny
    ln = np.linspace(c_lo, c_hi, lines + 2)[1:-1].round().astype(int)
    n1 = (np.repeat(ln, ny) * nny + np.tile(np.arange(ny), ln.size)).astype(np.int64)
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

    ms = time_steps(torch, D, step, steps)
    unique = nx * ny + lines * world * ny
    alg = sum(coo[m].v.numel() * 8 + csr[m].numel() * 8 for m in names)
    # correctness bit: sum of the assembled CSR entries of all ranks against the sum of the COO row slabs of the nodes
    # each rank owns: the same global sum when halo and row cuts are right
    loc = []
    for m in names:
        sm = 0.
        for g, b in enumerate(bs):
            own = (b.conn >= lo) & (b.conn < hi)         # row slabs of the nodes this rank owns
            off = plans[m].coo_offsets[g]
            sm += float(coo[m].v[off:off + b.ne * b.sizes[m]].view(b.ne, b.nn, -1)[own].sum())
        loc += [float(csr[m].sum()), float(csr[m].abs().sum()), sm]
    tot = D.reduce(loc)
    worst = max(abs(tot[3 * i] - tot[3 * i + 2]) / tot[3 * i + 1] for i in range(3))
    out = {"config": "5: stiffened panel %d x %d Quad4 + %d BeamC in one matrix each, KC0+KG+M + update_fint, %s"
                     % (nx, ny, lines * world * ny, "one GPU's share (1/8) of the 16M mesh" if world == 1 else
                        "row-sharded over %d GPUs" % world),
           "elements": int(unique), "elements_evaluated_per_gpu": int(bs[0].ne + bs[1].ne), "ms_per_step": ms,
           "elements_per_s": unique / ms * 1e3, "achieved_gbs_per_gpu": alg / ms / 1e6, "frac": alg / ms / 1e6 / peak,
           "fused_quad_share": not getattr(plan, "_fused_unsupported", False),
           "parity_csr_vs_coo": worst, "parity_ok": bool(worst <= 1e-11)}
    del plan, plans, coo, csr, fint, bs
    torch.cuda.empty_cache()
    return out


def solve_e2e(torch, meshes, dev, side):
    """User-level end to end on one GPU: host x in -> KC0 evaluated + assembled -> static problem solved by the native
    Jacobi-CG on the device -> only u back on the host (what a reference script does with its loop + tocsc + cg)."""
    import numpy as np
    from pyfe3d_b200.batch import AssemblyPlan
    from pyfe3d_b200.solve import plan_cg_native
    case, free, f, normal = meshes.static_case(side)
    b = meshes.batch_from_case(case, device=dev)
    nn = case["ndof"] // 6
    plan = AssemblyPlan("KC0", nn, [b])
    free_t = torch.as_tensor(free.astype(np.uint8)).to(dev)
    xh = torch.as_tensor(case["x"]).pin_memory()
    fh = torch.as_tensor(f).pin_memory()
    uh = torch.empty(6 * nn, dtype=torch.float64).pin_memory()

    def run():
        b.x.copy_(xh, non_blocking=True)
        ft = fh.to(dev, non_blocking=True)
        _, csr = plan.evaluate_assemble(KC0=True, write_coo=False)
        x, it, status, res, bn = plan_cg_native(plan, csr["KC0"], ft, free=free_t, rtol=1e-9, scaled_norm=True)
        uh.copy_(x, non_blocking=True)
        torch.cuda.synchronize()
        return it, status
    run()
    t0 = time.perf_counter()
    it, status = run()
    dt = time.perf_counter() - t0
    u = uh.numpy()
    w = u[0::6] * normal[0] + u[1::6] * normal[1] + u[2::6] * normal[2]
    return {"workload": "static solve, %dx%d Quad4 of the workload mesh, edges clamped, uniform normal load; "
                        "diagonally scaled CG to 1e-9 (scaled norm)" % (side, side),
            "elements": side * side, "dofs": 6 * nn, "seconds": dt, "cg_iterations": int(it), "cg_status": int(status),
            "w_max": float(np.abs(w).max()), "h2d_bytes": int(xh.numel() * 8 + fh.numel() * 8),
            "d2h_bytes": int(uh.numel() * 8)}


def sharded_cg(torch, dist, meshes, rank, world, dev, side, k1=64, k2=256):
    """N > 1: the one downstream exchange step of the path -- Jacobi-CG on KC0 of the side x side plate with the rows
    sharded over the ranks (SURVEY 8(e) + 8(f) rank 1): fused assembly of the own row block, then
    ``plan_cg_solve(group=WORLD)``: block SpMV on the own rows, point-to-point halo exchange of the search direction
    (grouped NCCL send / recv), two small all_reduces per iteration.  ms per iteration = slope between two iteration
    counts, device-timed, max over ranks; the same with an all_gather of the whole direction beside it."""
    import numpy as np
    from pyfe3d_b200 import sharding
    from pyfe3d_b200.batch import AssemblyPlan
    from pyfe3d_b200.solve import plan_cg_solve
    case, free, f, normal = meshes.static_case(side)
    n = case["ndof"]
    sub = sharding.shard_case(case, rank, world)
    b = meshes.batch_from_case(sub, device=dev)
    plan = AssemblyPlan("KC0", n // 6, [b], node_range=sub["owned_nodes"])
    _, csr = plan.evaluate_assemble(KC0=True, write_coo=False)
    vals = csr["KC0"]
    free_t = torch.as_tensor(free.astype(np.uint8)).to(dev)
    ft = torch.as_tensor(f).to(dev)

    def timed(iters):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x, info = plan_cg_solve(plan, vals, ft, free=free_t, rtol=0., maxiter=iters, group=dist.group.WORLD)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t), x

    out = {"what": "row-sharded Jacobi-CG on KC0 of the %dx%d plate (%d dofs, clamped edges): ms per iteration over %d ranks"
                   % (side, side, n, world), "rows_per_rank": plan.nrows, "nnz_per_rank": plan.nnz}
    xs = {}
    for mode in ("halo", "all_gather"):
        os.environ["PF3_CG_EXCHANGE"] = mode
        timed(32)
        t1, _ = timed(k1)
        t2, xs[mode] = timed(k2)
        out["ms_per_iteration_" + ("halo_p2p" if mode == "halo" else "all_gather")] = (t2 - t1) / (k2 - k1)
    os.environ["PF3_CG_EXCHANGE"] = "halo"
    # correctness bit: the iterates do not depend on HOW the direction reaches the other ranks -- after k2 iterations the
    # solution of the halo exchange must equal the one of the all_gather (a wrong or missing halo entry changes it at once)
    lo, hi = 6 * plan.node_begin, 6 * plan.node_end
    d = torch.stack([(xs["halo"][lo:hi] - xs["all_gather"][lo:hi]).abs().max(), xs["all_gather"][lo:hi].abs().max()])
    dist.all_reduce(d, op=dist.ReduceOp.MAX)
    out["halo_vs_all_gather_rel_diff_after_%d_iterations" % k2] = float(d[0] / d[1])
    out["ok"] = bool(float(d[0]) <= 1e-12 * float(d[1]))
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pyfe3d_b200 import meshes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D = Dist(torch, dist, dev, world)
    side = args.side
    peak, peak_src = measured_peak()

    # ---- headline: weak scaling, side x side elements per GPU
    weak = measure_plate(torch, D, meshes, side * world, side, float(world), rank, world, dev, args, full=True)
    # ---- the metric's own size cut into N strips
    strong = None
    if world > 1 and args.strong and side % world == 0:
        s = measure_plate(torch, D, meshes, side, side, 1.0, rank, world, dev, args, full=False)
        strong = {"elements_total": s["ne_total"], "elements_evaluated_per_gpu": s["ne_local"],
                  "ms_per_step": s["ms_step"], "value": s["ne_total"] / (s["ms_step"] * 1e-3), "unit": UNIT,
                  "parity_ok": s["parity"]["ok"], "kc0_abs_per_element": s["parity"]["kc0_abs_per_element"]}
    elif world == 1:
        strong = {"elements_total": weak["ne_total"], "elements_evaluated_per_gpu": weak["ne_local"],
                  "ms_per_step": weak["ms_step"], "value": weak["ne_total"] / (weak["ms_step"] * 1e-3), "unit": UNIT,
                  "note": "N=1: identical to the headline"}
    # ---- the other BASELINE configurations under the same clock
    others = []
    if args.others:
        try:
            f = args.others_scale
            if world == 1:
                others.append(other_config(torch, D, meshes, "2: BeamC curved cantilever, 100k elements, KC0 + M(mtype 0)",
                                           meshes.arc_beamc(int(100001 * f)), ("KC0", "M0"), peak, args.steps))
                others.append(other_config(torch, D, meshes, "3: Quad4R cylinder 1760x571 (1.0M), material axes, KC0 + KG_given_stress",
                                           meshes.cylinder_quad4r(int(1760 * f ** 0.5), int(571 * f ** 0.5)),
                                           ("KC0", "KGs"), peak, args.steps))
                others.append(other_config(torch, D, meshes, "4: Tria3R distorted plate 1415x1415x2 (4.0M), KC0 + M(mtype 1)",
                                           meshes.plate_tria3r(int(1415 * f ** 0.5), int(1415 * f ** 0.5)),
                                           ("KC0", "M1"), peak, args.steps))
            others.append(config5(torch, D, meshes, rank, world, dev, max(2, int(496 * f)), max(2, int(3968 * f ** 0.5)) if f < 1 else 3968,
                                  8, peak, args.steps))
        except Exception as exc:  # reporting only: the headline stands without it
            others.append({"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])})
    # ---- the downstream exchange step: row-sharded CG with a point-to-point halo exchange (N > 1)
    cg = None
    if world > 1 and args.cg_side > 0:
        try:
            cg = sharded_cg(torch, dist, meshes, rank, world, dev, args.cg_side)
        except Exception as exc:  # reporting only
            cg = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    # ---- user-level end to end (one GPU)
    solve = None
    if rank == 0 and args.solve_side > 0:
        try:
            solve = solve_e2e(torch, meshes, dev, args.solve_side)
        except Exception as exc:
            solve = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    D.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    fused = weak["fused"]
    kern, ne_local, ms_step = weak["kern"], weak["ne_local"], weak["ms_step"]
    if fused:
        names = ["quad_record_kernel + quad_fused_kernel<QUAD4> (COO values of KC0,KG,M + their CSR values; 2 launches)"]
        alg = [BYTES_PATH * ne_local]
        kern = kern[:1]
    else:
        names = ["quad_eval_kernel<QUAD4> (KC0+KG+M values)", "k_assemble (KC0)", "k_assemble (KG)", "k_assemble (M)"]
        alg = [BYTES_EVAL * ne_local,
               (576 * 8) * ne_local + 324 * 8 * ne_local, (144 * 8) * ne_local + 81 * 8 * ne_local,
               (480 * 8) * ne_local + 270 * 8 * ne_local]
    dom = int(np.argmax(kern))
    achieved = alg[dom] / (kern[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(alg[dom]),
                "kernel_ms": {n: float(k) for n, k in zip(names, kern)},
                "kernel_gbs": {n: float(a / (k * 1e-3) / 1e9) for n, a, k in zip(names, alg, kern)},
                "path_frac": BYTES_PATH * (ne_local / (ms_step * 1e-3)) / 1e9 / peak,
                "path_bytes_per_element": BYTES_PATH}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            t = json.load(f).get("fused" if fused else "twopass")
        if t and t.get("elements") == ne_local:
            roofline["traffic"] = t["dram_bytes_per_launch"]
            roofline["traffic_source"] = t.get("source")

    cpu = None
    if world == 1 and args.cpu_side > 0:
        try:
            from oracle import ref_loop
            if ref_loop.available():
                from oracle import cpu_bench
                rb = cpu_bench.ReferenceBench(side=args.cpu_side)
                rb.step()
                dt = rb.step()
                rb.close()
                cpu = {"value": rb.ne / dt, "unit": UNIT, "cores": rb.nproc, "kind": "reference", "sample": rb.describe(),
                       "element_loop_s": rb.loop_s, "scipy_tocsr_s": rb.tocsr_s}
                # SURVEY 8(d) / BASELINE.md: the same loop on ONE core (a smaller sample of the same mesh)
                r1 = cpu_bench.ReferenceBench(side=args.cpu_side_1core, nproc=1)
                dt1 = r1.step()
                r1.close()
                cpu["one_core"] = {"value": r1.ne / dt1, "unit": UNIT, "cores": 1, "sample": r1.describe(),
                                   "element_loop_s": r1.loop_s, "scipy_tocsr_s": r1.tocsr_s}
                if solve is not None and "seconds" in solve:
                    dts, its, wmax, info = cpu_bench.reference_static_solve(args.solve_side)
                    solve["reference_seconds"] = dts
                    solve["reference_cg_iterations"] = its
                    solve["reference_w_max"] = wmax
                    solve["w_max_rel_diff"] = abs(wmax - solve["w_max"]) / abs(wmax)
                    solve["speedup_vs_reference"] = dts / solve["seconds"]
                    solve["reference_what"] = ("compiled reference on 1 host core: Cython loop (KC0) + coo_matrix.tocsc + "
                                               "K[bu,:][:,bu] + diagonally scaled scipy cg, same mesh and tolerance")
        except Exception as exc:  # the CPU leg is reporting only
            cpu = {"value": None, "unit": UNIT, "error": str(exc)[:200]}

    value = weak["ne_total"] / (ms_step * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(side), "l2": L2_NOTE},   # identical in both arms; the rest: details
            "details": {"elements_total": weak["ne_total"], "elements_evaluated_per_gpu": ne_local,
                       "halo": "row-ownership strips, halo elements duplicated, no collective on the data path",
                       "output_gb_per_step_per_gpu": (BYTES_ASM_READ + BYTES_CSR) * ne_local / 1e9,
                       "symbolic_plan_s": weak["symbolic_s"], "indices": "values only (update_*v_only=1); plan built once",
                       "path": args.path, "strong": strong, "others": others, "solve_e2e": solve, "sharded_cg": cg},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": weak.get("e2e"), "gpu_launches": weak["launches"],
            "clocks": weak["clocks"], "parity": weak["parity"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


class numa_local:
    """Context manager: run the enclosed host allocations on the CPUs of the NUMA node the GPU hangs off
    (/sys/bus/pci/devices/<bus id>/numa_node), then restore the affinity.  Best effort: a no-op where /sys does not
    say (single-socket hosts, restricted containers).  Enters as a dict describing what was done."""

    def __init__(self, local):
        self.local, self.saved, self.info = local, None, {"node": None}

    def __enter__(self):
        try:
            import torch
            p = torch.cuda.get_device_properties(self.local)
            bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
            if node < 0:
                return self.info
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            allowed = os.sched_getaffinity(0)
            use = cpus & allowed
            if use:
                self.saved = allowed
                os.sched_setaffinity(0, use)
                self.info.update(node=node, cpus=len(use), gpu=bdf)
        except Exception as exc:   # noqa: BLE001 - diagnostics only
            self.info["note"] = str(exc)[:80]
        return self.info

    def __exit__(self, *a):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=2000, help="elements per side per GPU (2000 -> 4.0M Quad4)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--path", default="fused", choices=["fused", "twopass"],
                    help="fused: one node-centric kernel writes COO + CSR; twopass: eval kernel then 3 assemblies")
    ap.add_argument("--cpu-side", type=int, default=448,
                    help="side of the sub-plate the CPU arm evaluates (448^2 = 200 704 elements: BASELINE.md section 3)")
    ap.add_argument("--cpu-side-1core", type=int, default=448, help="side of the sub-plate of the one-core CPU figure")
    ap.add_argument("--e2e-upper", type=int, default=1, help="also time the end-to-end step returning one triangle")
    ap.add_argument("--strong", type=int, default=1, help="N>1: also time the metric's own size cut into N strips")
    ap.add_argument("--others", type=int, default=1, help="also time BASELINE configs 2-5 (config.others)")
    ap.add_argument("--others-scale", type=float, default=1.0, help="shrink the other configs (smoke tests)")
    ap.add_argument("--cg-side", type=int, default=2000, help="N>1: side of the plate of the row-sharded CG leg (0: skip)")
    ap.add_argument("--solve-side", type=int, default=32, help="side of the static-solve end-to-end workload (0: skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
