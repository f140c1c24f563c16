mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -15
python scripts/bench_configs.py 2>&1 | tee gpurun_out/configs.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    if 'kernel_ms' in d: print('   ', d['config'][:28], {k: round(v, 3) for k, v in d['kernel_ms'].items()})
    else: print('   ', d['config'][:44], d['path'], 'ms %.3f frac %.3f' % (d['ms_per_step'], d['frac_of_6538.9']))
"
