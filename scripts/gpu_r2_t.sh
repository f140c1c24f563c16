# round 2, visit T: full suite + default bench after the per-element-constants change
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest.txt 2>&1; tail -4 gpurun_out/r2t_pytest.txt
timeout 900 python bench.py --steps 20 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; tail -2 gpurun_out/r2t_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2t_bench.json").read().strip().splitlines()[-1])
print("value %.1fM ms %.3f frac %.4f launches %d" % (d["value"] / 1e6, d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"]))
print("parity ok", d["parity"]["ok"], "e2e", {k: (v if k != "symmetric_upper" else v["value"]) for k, v in d["e2e"].items() if k != "what"})
for o in d["details"]["others"]:
    print("other", {k: v for k, v in o.items() if k in ("config", "ms_per_step", "frac", "parity_ok")})
print("solve", {k: v for k, v in d["details"]["solve_e2e"].items() if k in ("seconds", "cg_iterations", "reference_seconds", "speedup_vs_reference", "w_max_rel_diff")})
PY
python scripts/bench_configs.py --kinds 2>&1 | head -3
