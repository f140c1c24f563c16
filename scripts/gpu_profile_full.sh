# full-size (4M Quad4) evidence: launch list of the step kernels + ncu full capture of K1/K2
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'quad_|k_assemble|k_fill|k_node' -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'quad_fused|quad_record' -s 6 -c 2 -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out
