# ncu evidence: full capture of the dominant kernels (SIDE small so replays are quick)
mkdir -p gpurun_out
SIDE=${SIDE:-700}
PATHSEL=${PATHSEL:-fused}
ncu --set full --clock-control none --import-source on -k regex:'quad_fused|quad_eval|k_assemble' -s 4 -c 2 -o gpurun_out/prof \
    python bench.py --side $SIDE --steps 2 --warmup 3 --e2e-steps 0 --cpu-side 0 --path $PATHSEL > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out
