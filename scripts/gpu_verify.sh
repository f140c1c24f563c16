# round-end rehearsal: the GPU parity suite, smoke(), our bench arm and the reference arm, as the driver runs them
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/verify_pytest.log 2>&1; tail -25 gpurun_out/verify_pytest.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -5
( time python bench.py --impl reference ) > gpurun_out/verify_bench_ref.json 2> gpurun_out/verify_bench_ref.err; tail -c 600 gpurun_out/verify_bench_ref.json; tail -4 gpurun_out/verify_bench_ref.err
( time python bench.py ) > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; tail -c 2500 gpurun_out/verify_bench.json; tail -4 gpurun_out/verify_bench.err
