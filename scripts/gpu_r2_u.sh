# round 2, visit U (N GPUs): the bench under torchrun at the driver's largest size
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2u_bench_${N}gpu.json 2> gpurun_out/r2u_bench_${N}gpu.err; tail -3 gpurun_out/r2u_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2u_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value %.1fM ms %.3f frac %.4f" % (d["value"] / 1e6, d["ms_per_step"], d["roofline"]["frac"]))
print("parity", d["parity"])
print("e2e", {k: (v if k != "symmetric_upper" else v["value"]) for k, v in d["e2e"].items() if k != "what"})
print("strong", d["details"]["strong"])
for o in d["details"]["others"]:
    print("other", {k: v for k, v in o.items() if k in ("config", "ms_per_step", "elements_per_s", "frac", "parity_ok")})
print("plan_s", d["details"]["symbolic_plan_s"])
PY
