mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/bench_configs.py --fint 2>&1 | tail -2 | cut -c1-130
python scripts/bench_configs.py --kinds 2>&1 | tail -7 | cut -c1-330
python scripts/bench_configs.py --config2 2>&1 | tail -1 | cut -c1-330
python bench.py --steps 20 --warmup 3 --e2e-steps 0 --cpu-side 0 --others 0 --solve-side 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['parity']['ok'])"
