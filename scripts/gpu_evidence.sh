# round evidence: bench line (ours + reference arm), ncu launch list of the same command, ncu full capture of K1/K2
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 1500 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'quad_fused|quad_record' -s 6 -c 2 -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/prof_bench.log 2>&1
python scripts/gpu_calib.py > gpurun_out/calib.json 2>&1; cat gpurun_out/calib.json
ls -la gpurun_out | head -30
