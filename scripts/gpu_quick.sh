python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-side 0 --path fused 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused 4M', round(d['value']/1e6,1),'Mel/s', round(d['ms_per_step'],3),'ms', round(d['roofline']['path_frac'],3), d['clocks']['reasons'])"
