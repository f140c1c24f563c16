mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --side 700 --steps 3 --warmup 3 --e2e-steps 0 --cpu-side 0 --path twopass 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
python bench.py --side 700 --steps 3 --warmup 3 --e2e-steps 0 --cpu-side 0 --path fused 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['path_frac'])"
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 --path fused > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -c 1500 gpurun_out/bench_fused.json; tail -5 gpurun_out/bench_fused.err
