mkdir -p gpurun_out
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
: > gpurun_out/r2z_fr.txt
for r in 1 2; do
for name in default r2 r3; do
  if [ "$name" = "default" ]; then cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else cp pyfe3d_b200/lib/variants/$name/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
  python scripts/bench_configs.py --fint 2>&1 | tail -2 | grep -o '"config": "[^"]*"\|"ms_plan_gather": [0-9.]*' | paste - - | sed "s/^/$name $r /" >> gpurun_out/r2z_fr.txt
done
done
cp pyfe3d_b200/lib/variants/r3/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quad4" 2>&1 | tail -1
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
sort -k3,5 -k1,1 gpurun_out/r2z_fr.txt
