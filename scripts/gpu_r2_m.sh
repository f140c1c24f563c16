# round 2, visit M: update_fint element kernel at 3 CTAs/SM (spills) against 2 (A/B, interleaved)
mkdir -p gpurun_out
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
for r in 1 2 3; do
  for v in default fint3; do
    if [ "$v" = "default" ]; then cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else cp pyfe3d_b200/lib/variants/$v/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
    echo "== $v $r"; python scripts/bench_configs.py --fint 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d['config'], 'ms', round(d['ms_plan_gather'], 4))
    except Exception: pass"
  done
done | tee gpurun_out/r2m_fint_ab.txt
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
