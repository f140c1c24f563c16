# ncu evidence for the two-pass kernels of the other configurations (small sizes; shares and per-kernel counters)
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'tria_eval|k_assemble_slabs|line_eval|quad_eval' -s 6 -c 6 -o gpurun_out/prof_other \
    python scripts/bench_configs.py --small > gpurun_out/prof_other.log 2>&1
python scripts/bench_configs.py 2>&1 | tail -4 > gpurun_out/configs.jsonl
cat gpurun_out/configs.jsonl | cut -c1-260
