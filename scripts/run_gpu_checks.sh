set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for dp in 0.0 0.1; do
timeout 200 python scripts/step_bench.py --batch 256 --steps 30 --dropout $dp 2>&1 | tail -1
timeout 200 python scripts/step_bench.py --batch 512 --steps 20 --dropout $dp 2>&1 | tail -1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dropout.csv python scripts/step_bench.py --batch 256 --steps 1 --warmup 1 --no-graph > gpurun_out/ncu_step.log 2>&1
python scripts/kernel_shares.py gpurun_out/launches_dropout.csv > gpurun_out/kernel_shares_v6_dropout.json; head -60 gpurun_out/kernel_shares_v6_dropout.json
