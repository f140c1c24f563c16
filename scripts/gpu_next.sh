mkdir -p gpurun_out
python scripts/bench_configs.py --aero 2>&1 | tee gpurun_out/aero.jsonl | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'quad_|k_assemble|k_fill|k_node|k_plan' -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/launches_bench.log 2>&1
tail -5 gpurun_out/launches.csv | cut -c1-300
