# A/B of library variants on config 4 (Tria3R fused) and config 3 (Quad4R, two matrices): default vs pyfe3d_b200/lib/variants/*
mkdir -p gpurun_out
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
for round in 1 2; do
for d in default pyfe3d_b200/lib/variants/*/; do
  if [ "$d" = "default" ]; then name=default; cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else name=$(basename $d); cp $d/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
  for cfg in ${CFGS:---config4}; do
    python scripts/bench_configs.py $cfg | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-10s %-50s %.3f ms'%('$name', d['config'][:50], d['ms_per_step']))"
  done
done
done
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py tests/test_configs.py tests/test_scenarios.py -m gpu -x -q 2>&1 | tail -3
