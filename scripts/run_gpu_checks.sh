timeout 300 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -15
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_engine.py -m gpu -x -q 2>&1 | tail -5
for b in 256; do timeout 300 python scripts/step_bench.py --batch $b 2>&1 | tail -2; done
