mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.txt 2>&1; tail -4 gpurun_out/r2n_pytest.txt
python scripts/bench_configs.py --cg > gpurun_out/r02_cg.jsonl 2>&1; cat gpurun_out/r02_cg.jsonl
bash scripts/gpu_r2_evidence.sh
