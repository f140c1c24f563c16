mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --side 500 --steps 3 --warmup 3 --e2e-steps 1 --cpu-side 0 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
