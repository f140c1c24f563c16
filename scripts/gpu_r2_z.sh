mkdir -p gpurun_out
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
: > gpurun_out/r2z_t.txt
for r in 1 2 3; do
for name in default t1 t2 t8; do
  if [ "$name" = "default" ]; then cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else cp pyfe3d_b200/lib/variants/$name/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
  echo -n "$name $r " >> gpurun_out/r2z_t.txt
  python scripts/bench_configs.py --config4 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' >> gpurun_out/r2z_t.txt
done
done
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
sort gpurun_out/r2z_t.txt
python -m pytest tests/test_gpu_fused.py tests/test_gpu_benchmark_parity.py -m gpu -x -q 2>&1 | tail -2
python scripts/bench_configs.py --config3 2>&1 | tail -1 | cut -c1-60,150-330
