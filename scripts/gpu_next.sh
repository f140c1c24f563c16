mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/bench_configs.py --kinds 2>&1 | tee gpurun_out/kinds.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(d['config'][:40], 'ms %.3f frac %.3f fint %.3f' % (d['ms'], d['frac_of_6538.9'], d['update_fint_ms']))"
