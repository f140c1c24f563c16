# round 2, visit S: bulk copies issued with warp-uniform operands (no waterfall loop): parity + A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_benchmark_parity.py -m gpu -q -x > gpurun_out/r2s_pytest.txt 2>&1; tail -4 gpurun_out/r2s_pytest.txt
CHECK=0 STEPS=20 ROUNDS=3 BENCH_ARGS="--others 0 --solve-side 0" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2s_variants.txt
