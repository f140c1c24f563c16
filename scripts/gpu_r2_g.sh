# round 2, visit G: L2 evict-first stores + bulk L2 prefetch of the records in K2: A/B (interleaved), then the full suite
mkdir -p gpurun_out
CHECK=0 STEPS=20 ROUNDS=3 BENCH_ARGS="--others 0 --solve-side 0" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2g_variants.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2g_pytest.txt 2>&1; tail -8 gpurun_out/r2g_pytest.txt
