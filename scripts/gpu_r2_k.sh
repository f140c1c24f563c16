# round 2, visit K: full suite + default bench with the prefetch table generalised (triangles, config 3) 
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest.txt 2>&1; tail -4 gpurun_out/r2k_pytest.txt
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
print("value %.1fM ms %.3f frac %.4f launches %d" % (d["value"] / 1e6, d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"]))
print("parity", d["parity"])
print("e2e", {k: v for k, v in d["e2e"].items() if k != "what"})
for o in d["details"]["others"]:
    print("other", {k: v for k, v in o.items() if k in ("config", "ms_per_step", "frac", "parity_ok", "path")})
print("solve", d["details"]["solve_e2e"])
print("cpu", d["cpu_baseline"])
PY
