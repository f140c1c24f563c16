mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/preflight_bench.json 2> gpurun_out/preflight_bench.err; tail -1 gpurun_out/preflight_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/preflight_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"], d["parity"]["ok"], d["gpu_launches"], d["clocks"])
print([ (o.get("config","")[:12], o.get("ms_per_step"), o.get("frac")) for o in d["details"]["others"]])
PY
