# round 2 evidence: bench lines (both arms), ncu launch list of the bench command, ncu full captures of K1/K2 (4 M
# Quad4), of config 3 (quad_fused<QUAD4R,4>) and config 2 (line_eval + assembly), calibration
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv,noheader
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_line.json 2> gpurun_out/r02_bench_ref.err; tail -c 300 gpurun_out/r02_bench_reference_line.json; echo
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err; tail -3 gpurun_out/r02_bench.err; head -c 600 gpurun_out/r02_bench_line.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 --others 0 --solve-side 0 > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'quad_fused|quad_record' -s 6 -c 2 -o gpurun_out/r02_prof \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --cpu-side 0 --others 0 --solve-side 0 > gpurun_out/r02_prof_bench.log 2>&1
ncu --set full --clock-control none -k regex:'quad_fused|quad_record' -s 8 -c 2 -o gpurun_out/r02_prof_cfg3 \
    python scripts/bench_configs.py --config3 > gpurun_out/r02_prof_cfg3.log 2>&1
ncu --set full --clock-control none -k regex:'line_eval|k_assemble' -s 9 -c 3 -o gpurun_out/r02_prof_cfg2 \
    python scripts/bench_configs.py --config2 > gpurun_out/r02_prof_cfg2.log 2>&1
ncu --set full --clock-control none -k regex:'tria_fused|tria_record' -s 8 -c 2 -o gpurun_out/r02_prof_cfg4 \
    python scripts/bench_configs.py --config4 > gpurun_out/r02_prof_cfg4.log 2>&1
python scripts/gpu_calib.py > gpurun_out/r02_calib.json 2>&1; cat gpurun_out/r02_calib.json
python scripts/bench_configs.py --spmv > gpurun_out/r02_spmv.jsonl 2>&1; cat gpurun_out/r02_spmv.jsonl
python scripts/bench_configs.py --kinds > gpurun_out/r02_kinds.jsonl 2>&1; cat gpurun_out/r02_kinds.jsonl
ls -la gpurun_out | head -40
