set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_engine.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
timeout 200 python scripts/step_bench.py --batch 256 --steps 30 2>&1 | tail -1 | cut -c1-150
(cd gpurun_tmp/base && timeout 200 python scripts/step_bench.py --batch 256 --steps 30 2>&1 | tail -1 | cut -c1-150)
done
timeout 200 python scripts/step_bench.py --batch 512 --steps 20 2>&1 | tail -1 | cut -c1-150
(cd gpurun_tmp/base && timeout 200 python scripts/step_bench.py --batch 512 --steps 20 2>&1 | tail -1 | cut -c1-150)
timeout 200 python scripts/gemm_check.py --only perf_ 2>&1 | cut -c1-260
