mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
for v in default noblocks; do
    if [ "$v" = "default" ]; then cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else cp pyfe3d_b200/lib/variants/$v/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
    echo "== $v"
    python scripts/bench_configs.py --config2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        print(d['config'][:40], d.get('kernel_ms') or ('ms %.4f frac %.3f' % (d['ms_per_step'], d['frac_of_6538.9'])))
    except Exception: pass"
    python scripts/bench_configs.py --kinds 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        if any(k in d['config'] for k in ('beam', 'truss', 'spring')): print(d['config'][:36], 'ms %.4f frac %.3f' % (d['ms'], d['frac_of_6538.9']))
    except Exception: pass"
done | tee gpurun_out/r2y_line_ab.txt
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
