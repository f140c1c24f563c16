mkdir -p gpurun_out
N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])"
tail -2 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/bench_config5.py > gpurun_out/config5_n$N.json 2> gpurun_out/config5_n$N.err
tail -1 gpurun_out/config5_n$N.json | cut -c1-600; tail -2 gpurun_out/config5_n$N.err
