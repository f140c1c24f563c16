# round evidence (v9): new GPU tests, launch list of the bench command, ncu full capture of K1/K2 at 4M elements
mkdir -p gpurun_out
python -m pytest tests/test_gpu_reference_scripts.py -m gpu -x -q 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'quad_fused|quad_record' -s 6 -c 2 -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out | head -30
