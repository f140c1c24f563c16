python scripts/bench_configs.py 2>&1 | tail -6
