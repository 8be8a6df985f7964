set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_ops_gpu.py tests/test_engine.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
timeout 200 python scripts/step_bench.py --batch 256 --steps 30 2>&1 | tail -1 | cut -c1-150
(cd gpurun_tmp/base && timeout 200 python scripts/step_bench.py --batch 256 --steps 30 2>&1 | tail -1 | cut -c1-150)
done
timeout 200 python scripts/step_bench.py --batch 512 --steps 20 2>&1 | tail -1 | cut -c1-150
(cd gpurun_tmp/base && timeout 200 python scripts/step_bench.py --batch 512 --steps 20 2>&1 | tail -1 | cut -c1-150)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v10.csv python scripts/step_bench.py --batch 256 --steps 1 --warmup 1 --no-graph > gpurun_out/ncu_c.log 2>&1
python scripts/kernel_shares.py gpurun_out/launches_v10.csv > gpurun_out/kernel_shares_v10.json; grep -E '"kernel"|total_us' gpurun_out/kernel_shares_v10.json | head -20
