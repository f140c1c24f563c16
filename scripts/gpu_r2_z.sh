# round 2, visit Z: A/B on the headline and configs 3/4 (prev = the committed build)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2z_pytest.txt 2>&1; tail -3 gpurun_out/r2z_pytest.txt
CHECK=0 STEPS=20 ROUNDS=3 BENCH_ARGS="--others 0 --solve-side 0" bash scripts/gpu_variants.sh 2>&1 | tee gpurun_out/r2z_variants.txt
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
for r in 1 2; do
  for v in default prev; do
    if [ "$v" = "default" ]; then cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else cp pyfe3d_b200/lib/variants/$v/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
    echo "== $v $r"
    for c in --config3 --config4; do python scripts/bench_configs.py $c 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d['config'][:40], 'ms', round(d['ms_per_step'], 4))
    except Exception: pass"; done
  done
done | tee gpurun_out/r2z_cfg_ab.txt
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
