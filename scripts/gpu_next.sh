mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k host_buffer 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 900 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
