# round 2, visit B: compact 32-double records + in-warp record dedup; ablations of the record fetch
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2b_pytest.txt 2>&1; tail -12 gpurun_out/r2b_pytest.txt
CHECK=0 STEPS=10 bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2b_variants.txt
