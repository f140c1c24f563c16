python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 --path twopass 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('twopass 4M', round(d['value']/1e6,1),'Mel/s', round(d['ms_per_step'],3),'ms', d['roofline']['kernel_ms'])"
python scripts/bench_configs.py 2>&1 | tail -4 | cut -c1-330
