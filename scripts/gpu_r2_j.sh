# round 2, visit J (2 GPUs): the bench under torchrun (weak + strong + config 5 + parity), the NCCL tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; tail -5 gpurun_out/r2j_bench_2gpu.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2j_bench_2gpu.json").read().strip().splitlines()[-1])
print("value %.1fM ms %.3f frac %.4f" % (d["value"] / 1e6, d["ms_per_step"], d["roofline"]["frac"]))
print("parity", d["parity"])
print("e2e", {k: v for k, v in d["e2e"].items() if k != "what"})
print("strong", d["details"]["strong"])
for o in d["details"]["others"]:
    print("other", o)
PY
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 | cut -c1-200
