# ncu evidence for the current round: launch list + full capture of the dominant kernels
mkdir -p gpurun_out
SIDE=${SIDE:-700}
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --side $SIDE --steps 2 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'quad_eval|k_assemble' -s 8 -c 4 -o gpurun_out/prof \
    python bench.py --side $SIDE --steps 2 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out
