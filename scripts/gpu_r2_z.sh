mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_line_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -1 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_line_2gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["parity"]["ok"], d["details"]["strong"]["ms_per_step"])
print(d["details"]["sharded_cg"])
PY
