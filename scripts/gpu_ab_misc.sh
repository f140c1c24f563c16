# A/B of the default library against pyfe3d_b200/lib/variants/* on the per-entry SpMV and the mixed-mesh path
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
for d in default pyfe3d_b200/lib/variants/*/; do
  if [ "$d" = "default" ]; then name=default; cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else name=$(basename $d); cp $d/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
  echo "== $name"
  python scripts/bench_configs.py --spmv | cut -c1-220
  python scripts/bench_configs.py --mixed | grep "fused quad share" | cut -c1-420
done
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
python -m pytest tests/test_gpu_spmv.py tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -3
