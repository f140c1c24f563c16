# round 2, visit P: producer mode (K1 inside K2, records made in L2): parity, then A/B against the two-launch build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py tests/test_gpu_benchmark_parity.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2p_pytest.txt 2>&1; tail -5 gpurun_out/r2p_pytest.txt
CHECK=0 STEPS=20 ROUNDS=3 BENCH_ARGS="--others 0 --solve-side 0" timeout 900 bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2p_variants.txt
