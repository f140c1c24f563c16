# round 2, visit O: occupancy variants of K2 with the cheaper reads; final micro-benchmark sweep; launch list; spmv test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmv.py -m gpu -q 2>&1 | tail -3
./scripts/micro/k2_stream_reads > gpurun_out/r02_k2_stream_reads.jsonl; wc -l gpurun_out/r02_k2_stream_reads.jsonl
CHECK=0 STEPS=20 ROUNDS=3 BENCH_ARGS="--others 0 --solve-side 0" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2o_variants.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 --others 0 --solve-side 0 > gpurun_out/r02_launches_bench.log 2>&1
tail -3 gpurun_out/r02_launches.csv | cut -c1-200
