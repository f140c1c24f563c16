# round 2, visit D: which reads cost the store stream its 1.3 ms (occupancy / dependency / prefetch sweeps); solver tests
mkdir -p gpurun_out
./scripts/micro/k2_stream_reads | tee gpurun_out/r2d_k2_stream_reads.jsonl
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_spmv.py tests/test_gpu_fused.py -m gpu -q --durations=5 > gpurun_out/r2d_pytest.txt 2>&1; tail -8 gpurun_out/r2d_pytest.txt
timeout 600 python bench_new.py --side 200 --others 0 --steps 3 --cpu-side 0 --solve-side 0 > gpurun_out/r2d_bench_new_small.json 2> gpurun_out/r2d_bench_new_small.err; tail -3 gpurun_out/r2d_bench_new_small.err; python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_new_small.json').read().strip().splitlines()[-1]); print(d['e2e'])"
