mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fint or finte or config5 or mixed or elements_api or reference" 2>&1 | tail -3
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
: > gpurun_out/r2z_pipe.txt
for r in 1 2 3; do
for name in default nopipe; do
  if [ "$name" = "default" ]; then cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else cp pyfe3d_b200/lib/variants/$name/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
  echo "== $name $r" >> gpurun_out/r2z_pipe.txt
  python scripts/bench_configs.py --fint 2>&1 | tail -2 | cut -c1-130 >> gpurun_out/r2z_pipe.txt
done
done
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
cat gpurun_out/r2z_pipe.txt
