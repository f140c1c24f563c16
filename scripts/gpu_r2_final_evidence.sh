# round 2, final evidence refresh: bench lines of both arms, launch list, ncu full capture of K1/K2 with source, config 3/4 captures
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_line.json 2> gpurun_out/r02_bench_ref.err; cut -c1-200 gpurun_out/r02_bench_reference_line.json
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_line_final.json 2> gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_bench_final.err; cut -c1-300 gpurun_out/r02_bench_line_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 --others 0 --solve-side 0 > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'quad_fused|quad_record' -s 6 -c 2 -o gpurun_out/r02_prof \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --cpu-side 0 --others 0 --solve-side 0 > gpurun_out/r02_prof_bench.log 2>&1
ncu --set full --clock-control none -k regex:quad_fused -s 4 -c 1 -o gpurun_out/r02_prof_cfg3 \
    python scripts/bench_configs.py --config3 > gpurun_out/r02_prof_cfg3.log 2>&1
ncu --set full --clock-control none -k regex:'tria_fused|tria_record' -s 8 -c 2 -o gpurun_out/r02_prof_cfg4 \
    python scripts/bench_configs.py --config4 > gpurun_out/r02_prof_cfg4.log 2>&1
python scripts/bench_configs.py --kinds > gpurun_out/r02_kinds.jsonl 2>&1
python scripts/bench_configs.py --fint > gpurun_out/r02_fint.jsonl 2>&1; cat gpurun_out/r02_fint.jsonl
ls -la gpurun_out/*.ncu-rep
