set -x
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_engine.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python scripts/gemm_check.py --only perf 2>&1 | cut -c1-420
timeout 300 python scripts/step_bench.py --batch 256 2>&1 | tail -1
timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 2>&1 | grep -v "Warning\|Writing\|Loading" | tail -3
