mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -4
for P in fused twopass; do
python bench.py --side 700 --steps 3 --warmup 3 --e2e-steps 0 --cpu-side 0 --path $P 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$P side700', round(d['value']/1e6,1),'Mel/s', round(d['ms_per_step'],3),'ms', round(d['roofline']['path_frac'],3))"
done
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 --path fused 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused 4M', round(d['value']/1e6,1),'Mel/s', round(d['ms_per_step'],3),'ms', round(d['roofline']['path_frac'],3), d['clocks'])"
