# round 2, visit C: full GPU suite; reads/compute micro-benchmark; new bench smoke + full default run
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2c_pytest.txt 2>&1; tail -15 gpurun_out/r2c_pytest.txt
./scripts/micro/k2_stream_reads | tee gpurun_out/r2c_k2_stream_reads.jsonl
CHECK=0 STEPS=10 bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2c_variants.txt
timeout 600 python bench_new.py --side 200 --others-scale 0.01 --steps 3 --cpu-side 32 --cpu-side-1core 16 --solve-side 12 > gpurun_out/r2c_bench_new_small.json 2> gpurun_out/r2c_bench_new_small.err; tail -3 gpurun_out/r2c_bench_new_small.err; head -c 2500 gpurun_out/r2c_bench_new_small.json; echo
timeout 900 python bench_new.py > gpurun_out/r2c_bench_new.json 2> gpurun_out/r2c_bench_new.err; tail -3 gpurun_out/r2c_bench_new.err; head -c 6000 gpurun_out/r2c_bench_new.json; echo
