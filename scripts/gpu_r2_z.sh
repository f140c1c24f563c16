python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
