python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k "host_buffer" 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --cpu-side 0 > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err; tail -3 gpurun_out/bench_pipe.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_pipe.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "el/s", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["clocks"])
PY
