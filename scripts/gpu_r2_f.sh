mkdir -p gpurun_out
./scripts/micro/k2_stream_reads | tee gpurun_out/r2f_k2_stream_reads.jsonl
