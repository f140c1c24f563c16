mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --e2e-steps 0 --cpu-side 0 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('north-star ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
python scripts/bench_configs.py --config3 2>&1 | tail -1 | cut -c1-250
python scripts/bench_configs.py --mixed 2>&1 | tail -4 | head -1 | cut -c1-250
