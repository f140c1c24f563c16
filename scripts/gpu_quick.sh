python scripts/bench_configs.py 2>&1 | grep kernel_ms
