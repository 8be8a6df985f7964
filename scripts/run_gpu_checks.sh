set -x
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_attention_gpu.py tests/test_engine.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do timeout 300 python scripts/step_bench.py --batch 256 --steps 40 2>&1 | tail -1 | cut -c1-200; done
DTB200_GEMM_CLUSTER=1 timeout 300 python scripts/step_bench.py --batch 256 --steps 40 2>&1 | tail -1 | cut -c1-200
DTB200_ATTN_NO_SMALL=1 timeout 300 python scripts/step_bench.py --batch 256 --steps 40 2>&1 | tail -1 | cut -c1-200
nvidia-smi --query-gpu=name,clocks.sm,power.draw,power.limit,temperature.gpu --format=csv
