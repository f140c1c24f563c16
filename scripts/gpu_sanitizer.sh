# compute-sanitizer over the small-size GPU tests: memcheck (out-of-bounds / misaligned accesses, TMA and cp.async included)
# on the fused, parity, solve, SpMV, aero and drop-in tests; racecheck (shared-memory hazards) on the fused kernels' tests.
# Runs against the -DPF3_MEMCHECK_BUILD library (scripts/build_memcheck_lib.sh; see prop_index in csrc/common.cuh for why).
mkdir -p gpurun_out
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
cp pyfe3d_b200/lib/variants/memcheck/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so
timeout 1500 compute-sanitizer --tool memcheck --show-backtrace no --print-limit 6 --error-exitcode 7 \
  python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_spmv.py tests/test_gpu_aero.py tests/test_gpu_elements_api.py -m gpu -q \
  > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep "Invalid\|at void\|at pf3\|ERROR SUMMARY\|passed\|failed\|^FAILED" gpurun_out/sanitizer_memcheck.log | sort | uniq -c | sort -rn | head -12
timeout 700 compute-sanitizer --tool memcheck --show-backtrace no --error-exitcode 7 \
  python -m pytest tests/test_gpu_solve.py -m gpu -q > gpurun_out/sanitizer_memcheck_solve.log 2>&1; echo "memcheck solve rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_solve.log
if [ "${RACE:-1}" = "1" ]; then
timeout 1000 compute-sanitizer --tool racecheck --show-backtrace no --error-exitcode 7 \
  python -m pytest tests/test_gpu_fused.py -m gpu -q -k "not 402" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
grep "hazard\|Race\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitizer_racecheck.log | sort | uniq -c | sort -rn | head -12
fi
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
