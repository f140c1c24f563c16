mkdir -p gpurun_out
python scripts/bench_configs.py --fint 2>&1 | tee gpurun_out/fint.jsonl | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:'tria_fused|tria_record' -s 8 -c 2 -o gpurun_out/prof_tria \
    python scripts/bench_configs.py --config4 > gpurun_out/prof_tria.log 2>&1
ncu --set full --clock-control none -k regex:'quad_eval' -s 2 -c 1 -o gpurun_out/prof_fint \
    python scripts/bench_configs.py --fint > gpurun_out/prof_fint.log 2>&1
ls -la gpurun_out/*.ncu-rep
