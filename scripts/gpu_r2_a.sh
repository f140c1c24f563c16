# round 2, visit A: K2 ablations + the dynamically scheduled variant; parity of the default and dyn builds
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv,noheader
CHECK=0 STEPS=10 bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2a_variants.txt
# parity: default build
timeout 2400 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2a_pytest_default.txt 2>&1; tail -30 gpurun_out/r2a_pytest_default.txt
# parity: dyn build
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
cp pyfe3d_b200/lib/variants/dyn/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2a_pytest_dyn.txt 2>&1; tail -5 gpurun_out/r2a_pytest_dyn.txt
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
