# bench every library variant under pyfe3d_b200/lib/variants on the GPU box (quick: device-resident metric only)
mkdir -p gpurun_out
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
for d in default pyfe3d_b200/lib/variants/*/; do
  if [ "$d" = "default" ]; then name=default; cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else name=$(basename $d); cp $d/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
  python bench.py --steps ${STEPS:-10} --warmup 3 --e2e-steps 0 --cpu-side 0 ${BENCH_ARGS:-} > gpurun_out/bench_var_$name.json 2> gpurun_out/bench_var_$name.err || tail -3 gpurun_out/bench_var_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bench_var_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("%-28s ms/step %.3f  el/s %.1fM  frac %.4f  sm %s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6, d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
if [ "${CHECK:-1}" = "1" ]; then python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4; fi
