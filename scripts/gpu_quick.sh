python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -25
