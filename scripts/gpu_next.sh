mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k tria 2>&1 | tail -5
CFG_ARGS=--config4 bash scripts/gpu_variants_cfg.sh
