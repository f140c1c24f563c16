mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'quad_fused' -s 4 -c 1 -o gpurun_out/r02_prof_cfg3 \
    python scripts/bench_configs.py --config3 > gpurun_out/r02_prof_cfg3.log 2>&1
ls -la gpurun_out/r02_prof_cfg3.ncu-rep
