# bench every library variant under pyfe3d_b200/lib/variants on the GPU box (device-resident metric only).
# ROUNDS passes over all variants, interleaved, so that box drift (clocks, power state: +-1.5 % between processes) hits
# every variant alike; prints min and median ms/step per variant.
mkdir -p gpurun_out
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
BENCH=${BENCH:-bench.py}
for r in $(seq 1 ${ROUNDS:-3}); do
for d in default pyfe3d_b200/lib/variants/*/; do
  if [ "$d" = "default" ]; then name=default; cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else name=$(basename $d); cp $d/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
  python $BENCH --steps ${STEPS:-10} --warmup 3 --e2e-steps 0 --cpu-side 0 ${BENCH_ARGS:-} > gpurun_out/bench_var_${name}_$r.json 2> gpurun_out/bench_var_$name.err || tail -3 gpurun_out/bench_var_$name.err
done
done
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
python - <<'PY'
import glob, json, os, statistics
res = {}
for f in sorted(glob.glob("gpurun_out/bench_var_*_[0-9]*.json")):
    name = os.path.basename(f)[len("bench_var_"):].rsplit("_", 1)[0]
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        res.setdefault(name, []).append(d["ms_per_step"])
    except Exception as e:
        res.setdefault(name, [])
for name, v in res.items():
    if v:
        print("%-24s min %.3f  median %.3f  runs %s" % (name, min(v), statistics.median(v), " ".join("%.3f" % t for t in v)))
    else:
        print("%-24s failed" % name)
PY
rm -f gpurun_out/bench_var_*_[0-9]*.json
if [ "${CHECK:-1}" = "1" ]; then python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4; fi
