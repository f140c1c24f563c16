mkdir -p gpurun_out
cp pyfe3d_b200/lib/libpyfe3d_b200.so /tmp/lib_default.so
: > gpurun_out/r2z_sub.txt
for r in 1 2; do
for name in default vol1700; do
  if [ "$name" = "default" ]; then cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so; else cp pyfe3d_b200/lib/variants/$name/libpyfe3d_b200.so pyfe3d_b200/lib/libpyfe3d_b200.so; fi
  python scripts/bench_configs.py --subsets 2>&1 | grep "KC0+M0" | grep -o '"config": "[^"]*"\|"ms_per_step": [0-9.]*' | paste - - | sed "s/^/$name $r /" >> gpurun_out/r2z_sub.txt
done
done
cp /tmp/lib_default.so pyfe3d_b200/lib/libpyfe3d_b200.so
sort -k3,6 -k1,1 gpurun_out/r2z_sub.txt
