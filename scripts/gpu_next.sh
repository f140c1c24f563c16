mkdir -p gpurun_out
./scripts/micro/dfma | tee gpurun_out/dfma.json
python scripts/bench_configs.py --kinds 2>&1 | tee gpurun_out/kinds.jsonl | cut -c1-330
