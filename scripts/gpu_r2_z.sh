mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 scripts/bench_cg_multi.py 2000 2>&1 | tail -2 | tee gpurun_out/r2z_cgmulti8.txt
