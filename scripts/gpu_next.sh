mkdir -p gpurun_out
python scripts/bench_config5.py 2>&1 | tail -2 | cut -c1-700
