# round 2, visit L: four-lanes-per-element fint kernel: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_elements_api.py tests/test_gpu_fused.py tests/test_configs.py tests/test_scenarios.py tests/test_gpu_reference_scripts.py -m gpu -q -x > gpurun_out/r2l_pytest.txt 2>&1; tail -4 gpurun_out/r2l_pytest.txt
python scripts/bench_configs.py --fint 2>&1 | tee gpurun_out/r2l_fint.jsonl
