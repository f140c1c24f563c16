mkdir -p gpurun_out
CHECK=0 STEPS=20 ROUNDS=3 BENCH_ARGS="--others 0 --solve-side 0" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2z_variants.txt
