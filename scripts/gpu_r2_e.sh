# round 2, visit E: bulk L2 prefetch in the store-stream micro; solver test
mkdir -p gpurun_out
./scripts/micro/k2_stream_reads | tee gpurun_out/r2e_k2_stream_reads.jsonl
timeout 900 python -m pytest tests/test_gpu_solve.py -m gpu -q > gpurun_out/r2e_pytest.txt 2>&1; tail -4 gpurun_out/r2e_pytest.txt
