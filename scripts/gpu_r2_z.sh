mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/preflight_bench.json 2> gpurun_out/preflight_bench.err; tail -1 gpurun_out/preflight_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/preflight_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"]["ok"], d["gpu_launches"], d["clocks"])
print([ (o.get("config","")[:12], o.get("ms_per_step"), o.get("frac")) for o in d["details"]["others"]])
PY
ncu --set full --clock-control none -k regex:quad_fused -s 4 -c 1 -o gpurun_out/r02_prof_cfg3 -f python scripts/bench_configs.py --config3 > gpurun_out/r02_prof_cfg3.log 2>&1; tail -1 gpurun_out/r02_prof_cfg3.log | cut -c1-80
RACE=0 bash scripts/gpu_sanitizer.sh
