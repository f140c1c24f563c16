python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k mixed 2>&1 | tail -3
python scripts/bench_configs.py --mixed 2>&1 | tail -4 | head -1 | cut -c1-600
