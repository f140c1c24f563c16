# full-size (4M Quad4) evidence: launch list of one bench run (plan creation skipped) + ncu full capture of K1/K2
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'quad_fused|quad_record' -s 6 -c 2 -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --cpu-side 0 > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out
torchrun --standalone --nnodes=1 --nproc-per-node=1 bench.py --gpus 1 --side 300 --steps 2 --warmup 3 --e2e-steps 1 --cpu-side 0 2>&1 | tail -1 | cut -c1-300
