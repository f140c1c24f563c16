# one GPU visit: parity tests, bench line, ncu full capture (with source) of K1/K2 at 4M elements
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "el/s", d["value"], "frac", d["roofline"]["frac"], d["clocks"])
PY
if [ "${PROFILE:-1}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:'quad_fused|quad_record' -s 6 -c 2 -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out/prof.ncu-rep
fi
