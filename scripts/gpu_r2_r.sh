# round 2, visit R: per-element constants (Quad4R hourglass, Quad4 thick flag) from K1: parity, A/B on the headline and config 3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2r_pytest.txt 2>&1; tail -4 gpurun_out/r2r_pytest.txt
CHECK=0 STEPS=20 ROUNDS=3 BENCH_ARGS="--others 0 --solve-side 0" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2r_variants.txt
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
for r in 1 2 3; do
  for v in default prev; do
    if [ "$v" = "default" ]; then cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else cp pyfe3d_b200/lib/variants/$v/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
    echo "== $v $r"; python scripts/bench_configs.py --config3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d['config'][:40], 'ms', round(d['ms_per_step'], 4))
    except Exception: pass"
  done
done | tee gpurun_out/r2r_cfg3_ab.txt
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
