# round 2, visit Q (2 GPUs): full GPU suite (incl. NCCL tests), bench at N=1 and N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.txt 2>&1; tail -4 gpurun_out/r2q_pytest.txt
timeout 900 python bench.py > gpurun_out/r2q_bench_1gpu.json 2> gpurun_out/r2q_bench_1gpu.err; tail -2 gpurun_out/r2q_bench_1gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2q_bench_2gpu.json 2> gpurun_out/r2q_bench_2gpu.err; tail -2 gpurun_out/r2q_bench_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r2q_bench_1gpu.json", "gpurun_out/r2q_bench_2gpu.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.1fM ms %.3f frac %.4f launches %d" % (d["value"] / 1e6, d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"]))
    print(" parity", d["parity"])
    print(" e2e", {k: (v if k != "symmetric_upper" else v["value"]) for k, v in d["e2e"].items() if k != "what"})
    print(" strong", d["details"]["strong"])
    for o in d["details"]["others"]:
        print(" other", {k: v for k, v in o.items() if k in ("config", "ms_per_step", "frac", "parity_ok", "parity_csr_vs_coo")})
PY
