mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmv.py -m gpu -x -q 2>&1 | tail -4
python scripts/bench_configs.py --spmv 2>&1 | tee gpurun_out/spmv.jsonl | cut -c1-330
