mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'tria_fused' -s 4 -c 1 -o gpurun_out/r02_prof_cfg4src \
    python scripts/bench_configs.py --config4 > gpurun_out/r02_prof_cfg4.log 2>&1
ls -la gpurun_out/r02_prof_cfg4src.ncu-rep
