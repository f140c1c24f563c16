# run scripts/bench_configs.py with every library variant
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
for d in default pyfe3d_b200/lib/variants/*/; do
  if [ "$d" = "default" ]; then name=default; cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else name=$(basename $d); cp $d/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
  echo "== $name"
  python scripts/bench_configs.py ${CFG_ARGS:-} 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    if 'kernel_ms' in d: print('   ', d['config'][:28], {k: round(v, 3) for k, v in d['kernel_ms'].items()})
    else: print('   ', d['config'][:44], d['path'], 'ms %.3f frac %.3f' % (d['ms_per_step'], d['frac_of_6538.9']))
"
done
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
