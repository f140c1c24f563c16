RACE=0 bash scripts/gpu_sanitizer.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
