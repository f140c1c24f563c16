# A/B of pf3_fill_indices (first-call cost): default library vs pyfe3d_b200/lib/variants/*; then the index parity tests
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
for d in default pyfe3d_b200/lib/variants/*/; do
  if [ "$d" = "default" ]; then name=default; cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else name=$(basename $d); cp $d/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
  echo "== $name"; python scripts/bench_configs.py --indices | cut -c1-200
done
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_elements_api.py tests/test_gpu_aero.py -m gpu -x -q 2>&1 | tail -3
